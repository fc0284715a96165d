"""Diagnostic (GPU): error of the values the final selection ranks, against the CPU oracle, per configuration.
    python scripts/selection_error.py"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from hoisdf_b200 import synthetic as syn  # noqa: E402
from hoisdf_b200.config import cfg  # noqa: E402
from hoisdf_b200.model import get_model  # noqa: E402
from oracle import hoisdf_oracle as O  # noqa: E402

dev = torch.device("cuda:0")
for arch in ("dexycb", "ho3d"):
    cfg.set_setting(arch)
    type(cfg).dataset = "ho3d"
    seed, B = 5, 2
    sd = syn.full_state_dict(seed, arch)
    model = get_model("test", mano_buffers=syn.mano_buffers(seed))
    model.load_state_dict(sd, strict=True)
    model = model.to(dev).eval()
    meta, pyr = syn.camera_meta(seed, B), syn.feature_pyramid(seed, B, arch)
    ocfg = O.default_cfg(num_samp_hand=600, num_samp_obj=200)
    d = lambda t: {k: v.to(dev) for k, v in t.items()}  # noqa: E731
    for kind, ck, bk, P in (("hand", "mano_root", "bbox_hand", 600), ("obj", "obj_center_cam", "bbox_obj", 200)):
        otaps = {}
        O.sdf_infer(dict(sd), pyr, meta[ck], meta["cam_intr"], meta[bk], 3.1, P, kind, ocfg, otaps)
        for tcp, final in ((False, "fma"), (True, "fma"), (True, "h3")):
            type(cfg).tc_projection, type(cfg).final_stage = tcp, final
            taps = {}
            with torch.no_grad():
                model.sdf_infer(d(pyr), d(meta)[ck], d(meta)["cam_intr"], d(meta)[bk], 3.1, P, kind, taps=taps)
            ex_sdf, ex_idx = taps["exact_sdf"].cpu().view(B, -1), taps["exact_index"].cpu().long().view(B, -1)
            worst, mism = 0.0, 0
            for b in range(B):
                pos = torch.searchsorted(otaps["cand_index"][b].contiguous(), ex_idx[b].contiguous())
                worst = max(worst, (ex_sdf[b] - otaps["cand_sdf"][b][pos]).abs().max().item())
                mism += int((taps["index"][b].cpu().long() != otaps["index"][b]).sum())
            print("%-7s %-4s projection=%-4s final=%-3s  max|sdf - oracle| = %.2e   index mismatches %d / %d   verified %s %s" % (
                arch, kind, "h3" if tcp else "fma", final, worst, mism, B * P, bool(taps.get("pre_verified", True)),
                bool(taps["screen_verified"])), flush=True)
