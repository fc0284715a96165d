"""Encoder self-attention of the training step at the BASELINE configs[3] shape (batch 64, 4 heads of 64, 800 tokens,
dropout 0.1 on the probabilities): CUDA-event times of AttentionFn forward / backward with the tensor-core backward
(hoisdf_attention_bwd) and with the batched fp32 FMA one, plus the algorithmic FLOP rate of each kernel.
    python scripts/attn_train_probe.py [batch] [tokens] [p_drop]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from hoisdf_b200 import autograd as A

B = int(sys.argv[1]) if len(sys.argv) > 1 else 64
L = int(sys.argv[2]) if len(sys.argv) > 2 else 800
P = float(sys.argv[3]) if len(sys.argv) > 3 else 0.1
H, d = 4, 256
dev = torch.device("cuda:0")
g = torch.Generator(device="cpu").manual_seed(0)
qkv = [torch.randn(B * L, d, generator=g).to(dev).requires_grad_() for _ in range(3)]
do = torch.randn(B * L, d, generator=g).to(dev)


def run(bwd_tc, iters=10):
    A._ATTN_BWD_TC = bwd_tc
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(3)]
    tf = tb = 0.0
    for it in range(iters + 3):
        for t in qkv:
            t.grad = None
        ev[0].record()
        out = A.AttentionFn.apply(*qkv, B, H, L, L, None, None, P)
        ev[1].record()
        out.backward(do)
        ev[2].record()
        torch.cuda.synchronize()
        if it >= 3:
            tf += ev[0].elapsed_time(ev[1]); tb += ev[1].elapsed_time(ev[2])
    return tf / iters, tb / iters, [t.grad.clone() for t in qkv]


flop_pair = 2.0 * B * H * L * L * 64          # one (L x L x 64) product over all (sample, head) pairs
f1, b1, g1 = run(True)
f0, b0, g0 = run(False)
print("attention B=%d H=%d L=%d p_drop=%.2f" % (B, H, L, P))
print("  forward  (tcgen05 flash, 2 products)            %.3f ms  %.1f TFLOP/s algorithmic" % (f1, 2 * flop_pair / f1 / 1e9))
print("  backward tensor cores (dQ + dK/dV kernels, 5 products + 2 recomputed)  %.3f ms  %.1f TFLOP/s algorithmic (5 products)"
      % (b1, 5 * flop_pair / b1 / 1e9))
print("  backward batched fp32 FMA (materialised S x S)  %.3f ms  %.1f TFLOP/s" % (b0, 5 * flop_pair / b0 / 1e9))
if P == 0.0:
    for a, b, n in zip(g1, g0, "qkv"):
        print("  d%s tensor-core vs fp32 FMA: max |diff| / max %.2e" % (n, float((a - b).abs().max() / b.abs().max())))
