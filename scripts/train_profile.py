"""Per-kernel time of ONE training step at the BASELINE configs[3] shape through torch.profiler (CUPTI; no ncu replay).
    python scripts/train_profile.py [batch]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from hoisdf_b200 import synthetic as syn
from hoisdf_b200.config import cfg
from hoisdf_b200.model import get_model
from hoisdf_b200.train import Trainer
B = int(sys.argv[1]) if len(sys.argv) > 1 else 64
dev = torch.device("cuda:0")
torch.backends.cudnn.benchmark = True
cfg.set_setting("ho3d"); type(cfg).num_samp_hand, type(cfg).num_samp_obj = 600, 200
model = get_model("train", mano_buffers=syn.mano_buffers(0)); model.load_state_dict(syn.full_state_dict(0, "ho3d"), strict=True)
model = model.to(dev).train()
tr = Trainer(model, lr=1e-4)
ins, tgt = syn.train_extras(100, B, 600, 200)
mv = lambda d: {k: v.to(dev) for k, v in d.items()}
batch = ({"img": syn.image_batch(100, B).to(dev), **mv(ins)}, mv(tgt), mv(syn.camera_meta(100, B)))
for _ in range(3): tr.step(*batch, epoch_cnt=0, batch_ratio=0.0)
torch.cuda.synchronize()
from torch.profiler import profile, ProfilerActivity
with profile(activities=[ProfilerActivity.CUDA]) as prof:
    tr.step(*batch, epoch_cnt=0, batch_ratio=0.0)
    torch.cuda.synchronize()
rows = sorted(((e.device_time_total if hasattr(e, "device_time_total") else e.cuda_time_total, e.count, e.key) for e in prof.key_averages()), reverse=True)
tot = sum(r[0] for r in rows)
print("one training step, batch %d: %.1f ms of kernel time in %d launches" % (B, tot / 1e3, sum(r[1] for r in rows)))
for t, n, k in rows[:40]:
    print("%8.2f ms %5.1f%% %5d  %s" % (t / 1e3, 100 * t / tot, n, k[:120]))
