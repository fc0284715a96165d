"""Developer tool: intermediate-gradient comparison of the training step (B200 kernels vs the oracle's autograd on CPU)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import torch
from hoisdf_b200 import synthetic as syn
from hoisdf_b200.config import cfg
from hoisdf_b200.model import get_model
from hoisdf_b200.train import total_loss
from oracle import hoisdf_oracle as O
from util import group_scales, param_group
torch.backends.cuda.matmul.allow_tf32 = False
torch.backends.cudnn.allow_tf32 = False
cuda = torch.device("cuda:0")
arch, seed, B, ph, po = "dexycb", 31, 2, 48, 16
cfg.set_setting(arch); type(cfg).dataset = "ho3d"
type(cfg).num_samp_hand, type(cfg).num_samp_obj, type(cfg).dropout, type(cfg).random_move_dist = ph, po, 0.0, [0.0] * 3
model = get_model("train", mano_buffers=syn.mano_buffers(seed))
sd = syn.full_state_dict(seed, arch)
model.load_state_dict(sd, strict=True)
model = model.to(cuda).train()
model.hand_sdf_decoder.dropout_prob = model.obj_sdf_decoder.dropout_prob = 0.0
mv = lambda d, dev: {k: v.clone().to(dev) for k, v in d.items()}
inputs, targets = syn.train_extras(seed, B, ph, po)
img, meta = syn.image_batch(seed, B), syn.camera_meta(seed, B)
model._train_debug = dbg = {}
out = model({"img": img.to(cuda), **mv(inputs, cuda)}, mv(targets, cuda), mv(meta, cuda), "train", 0, 0.0)
total, parts = total_loss(out)
total.backward()
# oracle
DT = torch.float64 if os.environ.get("ORACLE_F64", "1") == "1" else torch.float32
p = {k: (v.clone().to(DT) if v.is_floating_point() else v.clone()) for k, v in sd.items()}
names = [k for k, v in p.items() if v.is_floating_point() and "running_" not in k and "th_" not in k and "num_batches" not in k and "coord_change" not in k]
for n in names: p[n].requires_grad_(True)
taps = {}
cv = lambda d: {k: (v.clone().to(DT) if v.is_floating_point() else v.clone()) for k, v in d.items()}
oo = O.model_train(p, img.to(DT), cv(inputs), cv(targets), cv(meta), O.default_cfg(num_samp_hand=ph, num_samp_obj=po), arch, taps=taps)
keep = {}
for k in taps["pyramid"]:
    taps["pyramid"][k].retain_grad()
for k in ("hand_cls", "hand_off", "hand_fea", "obj_fea", "hand_transformer_in", "obj_transformer_in", "memory", "hand_encoder_out", "mano_pose6d", "mano_shape", "hand_joints"):
    taps[k].retain_grad(); keep[k] = taps[k]
ot, ol = O.train_total_loss(oo)
ot.backward()
def cmp(name, a, b):
    a, b = a.detach().cpu().double(), b.detach().cpu().double()
    print("%-28s value err %.3e (max %.3e)" % (name, float((a - b).abs().max()), float(b.abs().max())))
def cmpg(name, a, b):
    a, b = a.cpu().double(), b.double()
    print("%-28s GRAD err %.3e rel %.3e (max %.3e)" % (name, float((a - b).abs().max()), float((a - b).abs().max()) / max(float(b.abs().max()), 1e-30), float(b.abs().max())))
L = 3
cmp("hand_cls", dbg["hand_cls"].permute(0, 2, 1, 3), keep["hand_cls"]);      cmpg("hand_cls", dbg["hand_cls"].grad.permute(0, 2, 1, 3), keep["hand_cls"].grad)
cmp("hand_off", dbg["hand_off"].permute(0, 2, 1, 3), keep["hand_off"]);      cmpg("hand_off", dbg["hand_off"].grad.permute(0, 2, 1, 3), keep["hand_off"].grad)
cmp("hand_joints", dbg["hand_joints"], keep["hand_joints"]);                  cmpg("hand_joints", dbg["hand_joints"].grad, keep["hand_joints"].grad)
cmp("hand_fea", dbg["hand_fea"], keep["hand_fea"]);                           cmpg("hand_fea", dbg["hand_fea"].grad, keep["hand_fea"].grad)
cmp("obj_fea", dbg["obj_fea"], keep["obj_fea"]);                              cmpg("obj_fea", dbg["obj_fea"].grad, keep["obj_fea"].grad)
S = ph + po
cmp("hand_in", dbg["hand_in"].view(B, S, 256).transpose(0, 1), keep["hand_transformer_in"]); cmpg("hand_in", dbg["hand_in"].grad.view(B, S, 256).transpose(0, 1), keep["hand_transformer_in"].grad)
cmp("obj_in", dbg["obj_in"].view(B, S, 256).transpose(0, 1), keep["obj_transformer_in"]);   cmpg("obj_in", dbg["obj_in"].grad.view(B, S, 256).transpose(0, 1), keep["obj_transformer_in"].grad)
cmp("memory", dbg["memory"].view(B, S, 256).transpose(0, 1), keep["memory"]);               cmpg("memory", dbg["memory"].grad.view(B, S, 256).transpose(0, 1), keep["memory"].grad)
cmp("pose6d", dbg["pose6d"].permute(0, 2, 1, 3), keep["mano_pose6d"]);        cmpg("pose6d", dbg["pose6d"].grad.permute(0, 2, 1, 3), keep["mano_pose6d"].grad)
got = {n: q.grad for n, q in model.named_parameters() if q.grad is not None}
og = {n: p[n].grad for n in names if p[n].grad is not None}
sc = group_scales(og)
rows = sorted(((float((g.cpu().double() - og[n]).abs().max()) / sc[param_group(n)], n) for n, g in got.items()), reverse=True)
rows = [r for r in rows if not r[1].startswith(("backbone_net", "decoder_net"))]
pg = {k: v.grad for k, v in dbg["pyramid"].items()}
for e, n in rows[:25]:
    print("%.3e %s (own max %.3e, group max %.3e)" % (e, n, float(og[n].abs().max()), sc[param_group(n)]))

for k in pg:
    cmpg("pyramid " + k, pg[k], taps["pyramid"][k].grad)
# ReLU decisions that differ between the two evaluations (outputs of linear_transformerin: fea = relu(z))
for nm, a, b in (("hand_fea", dbg["hand_fea"], keep["hand_fea"]), ("obj_fea", dbg["obj_fea"], keep["obj_fea"])):
    a_, b_ = a.detach().cpu().double(), b.detach().double()
    flip = (a_ > 0) != (b_ > 0)
    print(nm, "ReLU decisions that differ:", int(flip.sum()), "of", flip.numel(),
          "values there (ours, oracle):", a_[flip].tolist()[:4], b_[flip].tolist()[:4],
          "d(fea) there:", a.grad.cpu()[flip].tolist()[:4], "max |d fea|", float(b.grad.abs().max()))
# per-group relative L2 error
import collections
num, den = collections.defaultdict(float), collections.defaultdict(float)
for n, g in got.items():
    num[param_group(n)] += float((g.cpu().double() - og[n].double()).pow(2).sum()); den[param_group(n)] += float(og[n].double().pow(2).sum())
print({k: "%.2e" % ((num[k] / den[k]) ** 0.5) for k in num})
