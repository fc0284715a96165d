"""Per-shape CUDA-event times of the FP16x3 GEMM launches of ONE training step (BASELINE configs[3] shape).
    python scripts/train_gemm_shapes.py [batch]"""
import os, sys, collections
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from hoisdf_b200 import ops, synthetic as syn
from hoisdf_b200.config import cfg
from hoisdf_b200.model import get_model
from hoisdf_b200.train import Trainer
B = int(sys.argv[1]) if len(sys.argv) > 1 else 64
dev = torch.device("cuda:0")
torch.backends.cudnn.benchmark = True
cfg.set_setting("ho3d"); type(cfg).num_samp_hand, type(cfg).num_samp_obj = 600, 200
type(cfg).native_sdf_infer = False
model = get_model("train", mano_buffers=syn.mano_buffers(0)); model.load_state_dict(syn.full_state_dict(0, "ho3d"), strict=True)
model = model.to(dev).train()
tr = Trainer(model, lr=1e-4)
ins, tgt = syn.train_extras(100, B, 600, 200)
mv = lambda d: {k: v.to(dev) for k, v in d.items()}
batch = ({"img": syn.image_batch(100, B).to(dev), **mv(ins)}, mv(tgt), mv(syn.camera_meta(100, B)))
for _ in range(2): tr.step(*batch, epoch_cnt=0, batch_ratio=0.0)
torch.cuda.synchronize()
ops.PROFILE = []
tr.step(*batch, epoch_cnt=0, batch_ratio=0.0)
torch.cuda.synchronize()
rec, ops.PROFILE = ops.PROFILE, None
agg = collections.OrderedDict()
for r in rec:
    if r[0] != "linear_h3":
        continue
    t = agg.setdefault(r[4], [0, 0.0, r[1]])
    t[0] += 1; t[1] += r[2].elapsed_time(r[3])
tot = sum(v[1] for v in agg.values())
print("FP16x3 GEMM launches of one training step (batch %d): %d launches, %.2f ms" % (B, sum(v[0] for v in agg.values()), tot))
for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1])[:45]:
    print("  %-44s x%-3d %7.3f ms  (%.3f each, %6.1f TFLOP/s)" % (k, v[0], v[1], v[1] / v[0], v[2] / (v[1] / v[0]) / 1e9))
