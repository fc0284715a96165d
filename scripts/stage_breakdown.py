"""Per-stage CUDA-event breakdown of one eval forward at the bench workload (developer tool, not a bench number)."""
import argparse, os, sys, collections, torch, warnings
warnings.filterwarnings("ignore")
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
ap = argparse.ArgumentParser()
ap.add_argument("--batch", type=int, default=32)
ap.add_argument("--ph", type=int, default=1536)
ap.add_argument("--po", type=int, default=512)
ap.add_argument("--arch", default="ho3d")
ap.add_argument("--iters", type=int, default=3)
ap.add_argument("--channels-last", action="store_true")
a = ap.parse_args()
from hoisdf_b200 import ops, synthetic as syn
from hoisdf_b200.config import cfg
from hoisdf_b200.model import get_model, PyramidContext
torch.backends.cudnn.allow_tf32 = False
torch.backends.cudnn.benchmark = True
dev = torch.device("cuda:0")
cfg.set_setting(a.arch); type(cfg).dataset = "ho3d"
type(cfg).num_samp_hand, type(cfg).num_samp_obj = a.ph, a.po
model = get_model("test", mano_buffers=syn.mano_buffers(0))
model.load_state_dict(syn.full_state_dict(0, a.arch)); model = model.to(dev).eval()
if a.channels_last: model.channels_last_()
B = a.batch
inputs = {"img": syn.image_batch(100, B).to(dev)}; targets = {k: v.to(dev) for k, v in syn.eval_targets(B).items()}
meta = {k: v.to(dev) for k, v in syn.camera_meta(100, B).items()}
times = collections.OrderedDict()
def wrap(obj, name, label):
    fn = getattr(obj, name)
    def w(*args, **kw):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); r = fn(*args, **kw); e1.record()
        times.setdefault(label, []).append((e0, e1)); return r
    setattr(obj, name, w)
wrap(model, "run_image_encoder", "image_encoder(resnet+unet)")
wrap(model, "sdf_infer", "sdf_infer"); wrap(model, "get_input_transformer", "point_features"); wrap(model, "sdf_forward", "sdf_forward(cross)")
wrap(model.hand_transformer, "forward_bm", "hand_transformer"); wrap(model.obj_transformer, "forward_bm", "obj_transformer")
wrap(model.mano_head, "forward_bm", "mano"); wrap(ops, "vote_joints", "vote")
import hoisdf_b200.model as MM
orig_g = PyramidContext.gmaps.fget
def g(self):
    if self._gmaps is None:
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); r = orig_g(self); e1.record(); times.setdefault("project_pyramid", []).append((e0, e1)); return r
    return orig_g(self)
PyramidContext.gmaps = property(g)
orig_h = MM.Model._head_rows
def h(*args, **kw):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); r = orig_h(*args, **kw); e1.record(); times.setdefault("heads", []).append((e0, e1)); return r
MM.Model._head_rows = staticmethod(h)
for it in range(a.iters + 2):
    if it == 2: times.clear()
    t0, t1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0.record(); out = model(inputs, targets, meta, "eval"); t1.record()
    times.setdefault("TOTAL", []).append((t0, t1))
torch.cuda.synchronize()
print("batch", B, "points", a.ph, a.po, "arch", a.arch, "N_f hand/obj mean",
      model.last_taps["hand"]["n_f"].double().mean().item(), model.last_taps["obj"]["n_f"].double().mean().item())
tot = sum(e0.elapsed_time(e1) for e0, e1 in times["TOTAL"]) / a.iters
for k, v in times.items():
    ms = sum(e0.elapsed_time(e1) for e0, e1 in v) / a.iters
    print("%-22s %9.3f ms  %5.1f%%  (%d calls/iter)" % (k, ms, 100 * ms / tot, len(v) // a.iters))
print("samples/s", B * 1000 / tot)
for kind in ("hand", "obj"):
    t = model.last_taps[kind]
    if "screen_gap" in t:
        print(kind, "screen verified", t["screen_verified"], "gap min %.3e" % float(t["screen_gap"].min()), "err %.3e" % float(t["screen_err"]))
