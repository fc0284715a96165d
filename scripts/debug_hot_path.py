import sys, os, torch, warnings
warnings.filterwarnings("ignore")
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from hoisdf_b200 import synthetic as syn
from hoisdf_b200.config import cfg
from hoisdf_b200.model import get_model
from oracle import hoisdf_oracle as O
dev = torch.device("cuda:0")
arch = "dexycb"
cfg.set_setting(arch); type(cfg).dataset = "ho3d"
type(cfg).num_samp_hand, type(cfg).num_samp_obj = 96, 40
seed, B = 5, 2
sd = syn.full_state_dict(seed, arch)
model = get_model("test", mano_buffers=syn.mano_buffers(seed)); model.load_state_dict(sd); model = model.to(dev).eval()
meta = syn.camera_meta(seed, B); pyr = syn.feature_pyramid(seed, B, arch)
out = model.hot_path({k: v.to(dev) for k, v in pyr.items()}, {k: v.to(dev) for k, v in meta.items()})
ot = {}
with torch.no_grad():
    oo = O.hot_path_eval(dict(sd), pyr, meta, O.default_cfg(num_samp_hand=96, num_samp_obj=40), ot)
t = model.last_taps
def r(a, b): 
    a, b = a.detach().cpu().double(), b.detach().cpu().double()
    return float((a-b).abs().max()/b.abs().max())
for k in ("hand_points","hand_sdf","hand_posenc","obj_points","obj_sdf","obj_posenc","hand_fea","obj_fea","hand_o_sdf","obj_h_sdf"):
    print(k, tuple(t[k].shape), r(t[k], ot[k]))
hi, oh = t["hand_transformer_in"].cpu(), ot["hand_transformer_in"].transpose(0,1)
for name, sl in (("xyz", slice(0,3)), ("pe", slice(3,33)), ("fea", slice(33,256))):
    for grp, ts in (("hand", slice(0,96)), ("obj_h", slice(96,136))):
        print("hand_in", name, grp, r(hi[:, ts, sl], oh[:, ts, sl]))
hi, oh = t["obj_transformer_in"].cpu(), ot["obj_transformer_in"].transpose(0,1)
for name, sl in (("xyz", slice(0,3)), ("pe", slice(3,33)), ("fea", slice(33,256))):
    for grp, ts in (("obj", slice(0,40)), ("hand_o", slice(40,136))):
        print("obj_in", name, grp, r(hi[:, ts, sl], oh[:, ts, sl]))
for k in ("hand_encoder_out","obj_encoder_out","hs"):
    print(k, r(t[k], ot[k].transpose(1,2)))
for k in ("hand_off","hand_cls","obj_rot","obj_trans"):
    print(k, r(t[k], ot[k].transpose(1,2)))
print("pose6d", r(t["mano_pose6d"], ot["mano_pose6d"].transpose(1,2)), "shape", r(t["mano_shape"], ot["mano_shape"]))
for k in oo: print(k, r(out[k], oo[k]))
