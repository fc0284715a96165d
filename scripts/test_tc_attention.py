import sys, os, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from hoisdf_b200 import ops
dev = torch.device("cuda:0")
torch.manual_seed(0)
def ref_attn(qkv, B, S, H, d, kv_valid=None):
    q, k, v = [t.double().view(B, S, H, 64).transpose(1, 2) for t in qkv.split(d, dim=2)]
    sc = q @ k.transpose(-1, -2) / 8.0
    if kv_valid is not None: sc[..., kv_valid:] = float("-inf")
    return (torch.softmax(sc, -1) @ v).transpose(1, 2).reshape(B, S, d)
for (B, S, kvv, scale) in [(1, 128, None, 1.0), (2, 64, None, 1.0), (2, 200, None, 1.0), (2, 801, None, 3.0), (1, 2048, None, 1.0), (2, 333, 150, 1.0), (3, 1000, None, 6.0)]:
    H, d = 4, 256
    qkv = (torch.randn(B, S, 3 * d, device=dev) * scale)
    q2 = qkv.view(B * S, 3 * d)
    ref = ref_attn(qkv, B, S, H, d, kvv)
    for tc in (True, False):
        ops.USE_TENSOR_CORES = tc
        out = torch.zeros(B * S, d, device=dev)
        ops.attention(q2, 3 * d, q2[:, d:], q2[:, 2 * d:], 3 * d, out, d, B, H, S, S, kv_valid=kvv)
        torch.cuda.synchronize()
        err = float((out.view(B, S, d).double() - ref).abs().max() / ref.abs().max())
        print("B=%d S=%d kv_valid=%s scale=%.0f %s rel.err %.2e" % (B, S, kvv, scale, "tc " if tc else "fma", err), flush=True)
B, S, H, d = 32, 2048, 4, 256
qkv = torch.randn(B, S, 3 * d, device=dev); q2 = qkv.view(B * S, 3 * d); out = torch.empty(B * S, d, device=dev)
for tc in (True, False):
    ops.USE_TENSOR_CORES = tc
    for _ in range(2): ops.attention(q2, 3 * d, q2[:, d:], q2[:, 2 * d:], 3 * d, out, d, B, H, S, S)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(5): ops.attention(q2, 3 * d, q2[:, d:], q2[:, 2 * d:], 3 * d, out, d, B, H, S, S)
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 5
    print("B=32 S=2048 %s: %.3f ms  %.1f TFLOP/s (4*S^2*d per sample)" % ("tc " if tc else "fma", ms, 4.0 * S * S * d * B / ms / 1e9), flush=True)
