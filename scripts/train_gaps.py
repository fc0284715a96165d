"""Where the GPU idles during one training step: gaps between consecutive kernels (torch.profiler / CUPTI timestamps),
aggregated by the kernel that ends the gap.   python scripts/train_gaps.py [batch]"""
import os, sys, collections
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
ev = sorted(((e.time_range.start, e.time_range.end, e.name) for e in prof.events()
             if e.device_type == torch.autograd.DeviceType.CUDA), key=lambda t: t[0])
busy = sum(e - s for s, e, _ in ev)
span = ev[-1][1] - ev[0][0]
gaps = collections.defaultdict(lambda: [0, 0.0])
end = ev[0][1]
timeline = []
for s, e, n in ev[1:]:
    if s > end:
        g = gaps[n[:70]]; g[0] += 1; g[1] += s - end
        timeline.append((s - ev[0][0], s - end, n[:60]))
    end = max(end, e)
idle = sum(v[1] for v in gaps.values())
print("one training step: span %.1f ms, kernels busy %.1f ms, idle %.1f ms in %d gaps" % (span / 1e3, busy / 1e3, idle / 1e3,
                                                                                      sum(v[0] for v in gaps.values())))
print("idle time by the kernel that ends the gap:")
for n, v in sorted(gaps.items(), key=lambda kv: -kv[1][1])[:25]:
    print("  %8.2f ms  x%-5d %s" % (v[1] / 1e3, v[0], n))
print("idle time per 10 ms of the step:")
bins = collections.Counter()
for t, g, _ in timeline:
    bins[int(t // 10000)] += g
print("  " + " ".join("%d:%.1f" % (b * 10, bins[b] / 1e3) for b in sorted(bins)))
print("largest single gaps:")
for t, g, n in sorted(timeline, key=lambda x: -x[1])[:12]:
    print("  at %7.2f ms  %7.3f ms before %s" % (t / 1e3, g / 1e3, n))
