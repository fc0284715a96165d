"""Which fp32 FMA GEMM calls (hoisdf_gemm_f32 / _batched: N <= 16 or K <= 16 layers, decoder attention) one training step makes,
with shapes and CUDA-event times.   python scripts/train_small_gemms.py [batch]"""
import os, sys, collections
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from hoisdf_b200 import synthetic as syn
from hoisdf_b200 import autograd as A
from hoisdf_b200._capi import lib
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
for _ in range(2): tr.step(*batch, epoch_cnt=0, batch_ratio=0.0)
torch.cuda.synchronize()
log = collections.OrderedDict()
real, real_b = lib.hoisdf_gemm_f32, lib.hoisdf_gemm_f32_batched
def timed(fn, key, args):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); r = fn(*args); e1.record(); torch.cuda.synchronize()
    t = log.setdefault(key, [0, 0.0]); t[0] += 1; t[1] += e0.elapsed_time(e1)
    return r
lib.hoisdf_gemm_f32 = lambda *a: timed(real, ("gemm", a[2], a[5], a[8], a[9], a[10]), a)
lib.hoisdf_gemm_f32_batched = lambda *a: timed(real_b, ("batched", a[2], a[7], a[14], a[15], a[16], a[19], a[20]), a)
tr.step(*batch, epoch_cnt=0, batch_ratio=0.0)
torch.cuda.synchronize()
tot = sum(v[1] for v in log.values())
print("fp32 FMA GEMM calls of one training step (batch %d): %.2f ms" % (B, tot))
for k, v in sorted(log.items(), key=lambda kv: -kv[1][1]):
    print("  %-60s x%-3d %.3f ms" % (k, v[0], v[1]))
