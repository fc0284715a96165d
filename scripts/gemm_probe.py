"""One FP16x3 GEMM shape, timed (developer tool; also the target of ncu captures).  python scripts/gemm_probe.py M N K [split] [reps]"""
import os, sys, torch, warnings
warnings.filterwarnings("ignore")
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from hoisdf_b200 import ops
dev = torch.device("cuda:0")
m, n, k = (int(a) for a in sys.argv[1:4])
split = len(sys.argv) > 4 and sys.argv[4] == "split"
reps = int(sys.argv[5]) if len(sys.argv) > 5 else 20
torch.manual_seed(0)
pw = ops.PackedLinearH3.pack(torch.randn(n, k, device=dev) * 0.05, torch.randn(n, device=dev))
xs = ops.split_rows(torch.randn(m, k, device=dev))
out = ops.SplitRows.empty(m, n, dev) if split else torch.empty(m, ops.round_up(n, 4), device=dev)[:, :n]
for _ in range(3):
    ops.linear_h3(xs, pw, 1, out=out)
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(reps):
    ops.linear_h3(xs, pw, 1, out=out)
e1.record()
torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / reps
print("M=%d N=%d K=%d %s: %.4f ms  %.1f TFLOP/s fp32-equivalent (%.1f executed)" % (m, n, k, "split" if split else "f32", ms,
      2.0 * m * n * k / ms / 1e9, 6.0 * m * n * k / ms / 1e9))
