"""Seed sweep of the point selection at the configs[1] shape (ho3d arch, 1536 + 512 points): for every seed the selected
lattice-index SETS of hoisdf_sdf_infer_fwd against the oracle's (upstream model.py:345-355 on fp32 CPU arithmetic), the
cascade's device-side verdict (gap / error) and the worst |sdf - oracle| of the values the final ranking used.
    python scripts/selection_sweep.py [n_seeds]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import torch
from hoisdf_b200 import synthetic as syn
from hoisdf_b200.config import cfg
from hoisdf_b200.model import get_model
from oracle import hoisdf_oracle as O
n_seeds = int(sys.argv[1]) if len(sys.argv) > 1 else 8
dev = torch.device("cuda:0")
cfg.set_setting("ho3d"); type(cfg).num_samp_hand, type(cfg).num_samp_obj = 1536, 512
B = 2
print("| seed | field | N_f per sample | selected sets identical | positions where the ORDER differs (near-ties of ~1e-8; the sets are what upstream consumes) | screening gap min | screening err | max abs(sdf - oracle) of ranked rows |")
print("|---|---|---|---|---|---|---|---|")
bad = 0
for seed in range(40, 40 + n_seeds):
    sd = syn.full_state_dict(seed, "ho3d")
    model = get_model("test", mano_buffers=syn.mano_buffers(seed)); model.load_state_dict(sd, strict=True); model = model.to(dev).eval()
    meta, pyr = syn.camera_meta(seed, B), syn.feature_pyramid(seed, B, "ho3d")
    to = lambda d: {k: v.to(dev) for k, v in d.items()}
    ocfg = O.default_cfg(num_samp_hand=1536, num_samp_obj=512)
    with torch.no_grad():
        ctx = model._ctx(to(pyr)); md = to(meta)
        for kind, ck, bk, P in (("hand", "mano_root", "bbox_hand", 1536), ("obj", "obj_center_cam", "bbox_obj", 512)):
            taps, otaps = {}, {}
            model.sdf_infer(ctx, md[ck], md["cam_intr"], md[bk], 3.1, P, kind, taps=taps)
            O.sdf_infer(dict(sd), pyr, meta[ck], meta["cam_intr"], meta[bk], 3.1, P, kind, ocfg, otaps)
            got, want = taps["index"].cpu().long(), otaps["index"]
            same = all(sorted(got[b].tolist()) == sorted(want[b].tolist()) for b in range(B))
            swaps = int((got != want).sum())
            ex_sdf, ex_idx = taps["exact_sdf"].cpu().view(B, -1), taps["exact_index"].cpu().long().view(B, -1)
            worst = 0.0
            for b in range(B):
                pos = torch.searchsorted(otaps["cand_index"][b].contiguous(), ex_idx[b].contiguous())
                worst = max(worst, float((ex_sdf[b] - otaps["cand_sdf"][b][pos]).abs().max()))
            bad += 0 if same else 1
            print("| %d | %s | %s | %s | %d | %.2e | %.2e | %.1e |" % (seed, kind, otaps["n_f"].tolist(), same, swaps,
                  float(taps["screen_gap"].min()), float(taps["screen_err"]), worst))
print("fields with a differing selected set: %d of %d" % (bad, 2 * n_seeds))
