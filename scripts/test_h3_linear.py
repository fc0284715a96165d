"""Developer check of the FP16x3 tensor-core Linear (csrc/linear_h3.cu): accuracy vs fp64 and speed vs 3xTF32."""
import ctypes, os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from hoisdf_b200 import ops
from hoisdf_b200._capi import lib
dev = torch.device("cuda:0")
torch.manual_seed(0)
lib.hoisdf_debug_h3_cluster.argtypes = [ctypes.c_int]


def run(m, n, k, act=0, res=False, split=False, cl=0, scale=1.0, batch=None):
    lib.hoisdf_debug_h3_cluster(cl)
    x = torch.randn(m, k, device=dev) * scale
    w = torch.randn(n, k, device=dev) * 0.05
    b = torch.randn(n, device=dev)
    pw = ops.PackedLinearH3.pack(w, b)
    xs = ops.split_rows(x)
    xj = xs.float()
    rt = float((xj - x).abs().max() / x.abs().max())
    r = torch.randn(m, ops.round_up(n, 4), device=dev)[:, :n] if res else None
    if batch:   # strided row groups: take the first `batch` rows of every group of 2*batch rows
        xb = torch.randn(m * 2, k, device=dev) * scale
        xs = ops.split_rows(xb)
        x = xb.view(-1, 2 * batch, k)[:, :batch].reshape(m, k)
        y = ops.linear_h3(xs, pw, act, residual=r, split_out=split, x_batch=(batch, 2 * batch * xs.ld), m=m)
    else:
        y = ops.linear_h3(xs, pw, act, residual=r, split_out=split)
    torch.cuda.synchronize()
    ref = x.double() @ w.double().T + b.double()
    if res: ref = ref + r.double()
    if act: ref = ref.relu()
    yf = y.float() if split else y
    err = float((yf.double() - ref).abs().max() / ref.abs().max())
    print("m=%d n=%d k=%d act=%d res=%d split=%d cl=%d batch=%s: rel.err %.2e (split round trip %.1e)"
          % (m, n, k, act, res, split, cl, batch, err, rt), flush=True)
    assert err < 4e-6 * max(1.0, k / 512), err
    return err


for cl in (1, 2, 4, 0):
    for shp in [(128, 256, 32), (128, 256, 64), (128, 16, 32), (1, 1, 8), (300, 512, 512), (1000, 223, 512),
                (257, 1024, 3968), (513, 768, 256), (64, 3, 256), (999, 512, 289), (130, 60, 256), (5000, 512, 519)]:
        run(*shp, cl=cl)
    run(1000, 512, 289, act=1, res=True, cl=cl)
    run(777, 256, 1024, act=0, res=True, cl=cl)
    run(3000, 512, 512, act=1, split=True, cl=cl)
    run(1000, 223, 512, act=1, split=True, cl=cl)
    run(2048, 256, 256, act=1, cl=cl, batch=128)
    run(1200, 60, 256, act=0, cl=cl, batch=100)
run(4096, 512, 512, scale=100.0)
run(4096, 512, 512, scale=1e-3)
lib.hoisdf_debug_h3_cluster(0)

# timing
for (m, n, k) in [(1 << 20, 512, 512), (1 << 20, 256, 512), (65536, 1024, 3968), (65536, 768, 256), (1 << 20, 512, 289),
                  (1 << 20, 223, 512)]:
    x = torch.randn(m, k, device=dev); w = torch.randn(n, k, device=dev) * 0.05; b = torch.randn(n, device=dev)
    pw3 = ops.PackedLinear.pack(torch.nn.functional.pad(w, (0, (-k) % 4)), b)
    xp = torch.nn.functional.pad(x, (0, (-k) % 4))
    out = torch.empty(m, n if n % 4 == 0 else ops.round_up(n, 4), device=dev)[:, :n]
    pwh = ops.PackedLinearH3.pack(w, b); xs = ops.split_rows(x); outs = ops.SplitRows.empty(m, n, dev)
    def bench(fn, tag):
        for _ in range(2): fn()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(5): fn()
        e1.record(); torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / 5
        print("m=%d n=%d k=%d %-22s: %.3f ms  %.1f TFLOP/s" % (m, n, k, tag, ms, 2.0 * m * n * k / ms / 1e9), flush=True)
    bench(lambda: ops.linear(xp, pw3, 1, out=out), "3xTF32")
    for cl in (1, 2, 4):
        lib.hoisdf_debug_h3_cluster(cl)
        bench(lambda: ops.linear_h3(xs, pwh, 1, out=out), "FP16x3 f32-out cl=%d" % cl)
        bench(lambda: ops.linear_h3(xs, pwh, 1, out=outs), "FP16x3 split-out cl=%d" % cl)
    lib.hoisdf_debug_h3_cluster(0)
