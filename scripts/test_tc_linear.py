import sys, os, torch, math, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from hoisdf_b200 import ops
dev = torch.device("cuda:0")
torch.manual_seed(0)
def run(m, n, k, act=0, res=False, ldx=None, check=True):
    ldx = ldx or k
    xb = torch.randn(m, ldx, device=dev); x = xb[:, :k]
    w = torch.randn(n, k, device=dev) * 0.05; b = torch.randn(n, device=dev)
    pw = ops.PackedLinear.pack(w, b)
    r = torch.randn(m, ops.round_up(n, 4), device=dev)[:, :n] if res else None
    out = torch.zeros(m, ops.round_up(n, 4), device=dev)[:, :n]
    ops.USE_TENSOR_CORES = True
    y = ops.linear(x, pw, act, out=out, residual=r)
    torch.cuda.synchronize()
    ref = x.double() @ w.double().T + b.double()
    if res: ref = ref + r.double()
    if act: ref = ref.relu()
    err = float((y.double() - ref).abs().max() / ref.abs().max())
    ops.USE_TENSOR_CORES = False
    y2 = ops.linear(x, pw, act, residual=r)
    torch.cuda.synchronize()
    err2 = float((y2.double() - ref).abs().max() / ref.abs().max())
    print("m=%d n=%d k=%d act=%d res=%d  tc rel.err %.2e   fma rel.err %.2e" % (m, n, k, act, res, err, err2), flush=True)
    return err
for shp in [(128, 256, 32), (128, 256, 64), (128, 16, 32), (1, 1, 4), (300, 512, 512), (1000, 223, 512), (257, 1024, 3968), (513, 768, 256), (64, 3, 256), (999, 512, 292), (130, 60, 256)]:
    run(*shp)
run(1000, 512, 292, act=1, res=True, ldx=516)
run(777, 256, 1024, act=0, res=True)
# timing
ops.USE_TENSOR_CORES = True
for (m, n, k) in [(1 << 20, 512, 512), (1 << 20, 256, 512), (65536, 1024, 3968), (65536, 768, 256), (1 << 20, 512, 292)]:
    x = torch.randn(m, k, device=dev); w = torch.randn(n, k, device=dev) * 0.05; b = torch.randn(n, device=dev)
    pw = ops.PackedLinear.pack(w, b); out = torch.empty(m, n, device=dev)
    for tc in (True, False):
        ops.USE_TENSOR_CORES = tc
        for _ in range(2): ops.linear(x, pw, 1, out=out)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(5): ops.linear(x, pw, 1, out=out)
        e1.record(); torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / 5
        print("m=%d n=%d k=%d %s: %.3f ms  %.1f TFLOP/s" % (m, n, k, "tc " if tc else "fma", ms, 2.0 * m * n * k / ms / 1e9), flush=True)
