"""Per-launch table of the FP16x3 GEMM launches of ONE bench step (CUDA events around every launch):
shape, ms, fp32-equivalent TFLOP/s.  python scripts/h3_launch_table.py [--batch 32]"""
import argparse
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402
from hoisdf_b200 import ops, synthetic as syn  # noqa: E402
from hoisdf_b200.config import cfg  # noqa: E402
from hoisdf_b200.model import get_model  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--batch", type=int, default=32)
args = ap.parse_args()
dev = torch.device("cuda:0")
torch.backends.cudnn.allow_tf32 = False
cfg.set_setting(bench.ARCH)
type(cfg).num_samp_hand, type(cfg).num_samp_obj = bench.P_HAND, bench.P_OBJ
model = get_model("test", mano_buffers=syn.mano_buffers(0))
model.load_state_dict(syn.full_state_dict(0, bench.ARCH), strict=True)
model = model.to(dev).eval()
inputs, targets, meta = bench.make_inputs(100, args.batch)
d = lambda t: {k: v.to(dev) for k, v in t.items()}  # noqa: E731
di, dt, dm = d(inputs), d(targets), d(meta)
for _ in range(3):
    model(di, dt, dm, "eval")
torch.cuda.synchronize()
ops.PROFILE = []
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
model(di, dt, dm, "eval")
e1.record()
torch.cuda.synchronize()
prof, ops.PROFILE = ops.PROFILE, None
print("step %.2f ms (with per-launch events)" % e0.elapsed_time(e1))
tot = {}
for i, p in enumerate(prof):
    ms = p[2].elapsed_time(p[3])
    tot.setdefault(p[0], [0.0, 0.0])
    tot[p[0]][0] += ms
    tot[p[0]][1] += p[1]
    print("%3d %-14s %8.3f ms %7.1f TF/s  %s" % (i, p[0], ms, p[1] / ms / 1e9, p[4] if len(p) > 4 else ""))
for k, (ms, fl) in tot.items():
    print("TOTAL %-14s %8.3f ms %7.1f TF/s" % (k, ms, fl / ms / 1e9))
for kind in ("hand", "obj"):
    t = model.last_taps[kind]
    msg = ["%s: single_pass=%s" % (kind, t.get("single_pass"))]
    for k in ("pre_err", "pre_gap", "screen_err", "screen_gap"):
        if k in t:
            msg.append("%s=%.3e" % (k, float(t[k].min())))
    for k in ("pre_verified", "screen_verified"):
        if k in t:
            msg.append("%s=%s" % (k, bool(t[k])))
    print(" ".join(msg))
