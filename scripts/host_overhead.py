"""Host enqueue time vs GPU time of one bench step (is the step launch-bound?).  python scripts/host_overhead.py"""
import os, sys, time
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402
from hoisdf_b200 import ops, synthetic as syn  # noqa: E402
from hoisdf_b200.config import cfg  # noqa: E402
from hoisdf_b200.model import get_model  # noqa: E402

dev = torch.device("cuda:0")
cfg.set_setting(bench.ARCH)
type(cfg).num_samp_hand, type(cfg).num_samp_obj = bench.P_HAND, bench.P_OBJ
model = get_model("test", mano_buffers=syn.mano_buffers(0))
model.load_state_dict(syn.full_state_dict(0, bench.ARCH), strict=True)
model = model.to(dev).eval()
if len(sys.argv) > 1 and sys.argv[1] == "graphs":
    model.enable_cuda_graphs()
inputs, targets, meta = bench.make_inputs(100, 32)
d = lambda t: {k: v.to(dev) for k, v in t.items()}  # noqa: E731
di, dt, dm = d(inputs), d(targets), d(meta)
for _ in range(4):
    model(di, dt, dm, "eval")
torch.cuda.synchronize()
host, gpu = [], []
for _ in range(8):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    e0.record()
    model(di, dt, dm, "eval")
    e1.record()
    host.append((time.perf_counter() - t0) * 1e3)
    torch.cuda.synchronize()
    gpu.append(e0.elapsed_time(e1))
print("host enqueue ms/step: min %.2f median %.2f   |   GPU ms/step (isolated): min %.2f median %.2f   | launches %d" % (
    min(host), sorted(host)[4], min(gpu), sorted(gpu)[4], ops.STATS["launches"] // 12))
