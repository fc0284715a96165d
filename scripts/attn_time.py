"""CUDA-event timing of the encoder self-attention kernel alone (B = 32, H = 4, S = 2048, d = 64)."""
import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from hoisdf_b200 import ops
dev = torch.device("cuda:0")
B, H, S, d = 32, 4, 2048, 256
qkv = torch.randn(B * S, 3 * d, device=dev)
out = ops.SplitRows.empty(B * S, d, dev)
for _ in range(3):
    ops.attention(qkv, 3 * d, qkv[:, d:], qkv[:, 2 * d:], 3 * d, out, d, B, H, S, S)
torch.cuda.synchronize()
e = [torch.cuda.Event(enable_timing=True) for _ in range(2)]
e[0].record()
for _ in range(10):
    ops.attention(qkv, 3 * d, qkv[:, d:], qkv[:, 2 * d:], 3 * d, out, d, B, H, S, S)
e[1].record()
torch.cuda.synchronize()
ms = e[0].elapsed_time(e[1]) / 10
print("attention incl. the 3 split kernels: %.3f ms  (%.1f TFLOP/s fp32-equivalent)" % (ms, 4.0 * S * S * 64 * H * B / ms / 1e9))
