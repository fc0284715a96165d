import sys, os, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from hoisdf_b200 import ops
dev = torch.device("cuda:0")
for (m, n, k) in [(1 << 20, 512, 512), (1 << 20, 256, 512), (1 << 20, 512, 292)]:
    x = torch.randn(m, k, device=dev); w = torch.randn(n, k, device=dev) * 0.05; b = torch.randn(n, device=dev)
    pw = ops.PackedLinear.pack(w, b); out = torch.empty(m, n, device=dev)
    ref = None
    for passes in (3, 1):
        for _ in range(2): ops.linear(x, pw, 1, out=out, passes=passes)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(5): ops.linear(x, pw, 1, out=out, passes=passes)
        e1.record(); torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / 5
        if ref is None: ref = out.clone()
        err = float((out - ref).abs().max() / ref.abs().max())
        print("m=%d n=%d k=%d passes=%d: %.3f ms  %.1f TFLOP/s  rel.diff vs 3-pass %.2e" % (m, n, k, passes, ms, 2.0 * m * n * k / ms / 1e9, err), flush=True)
