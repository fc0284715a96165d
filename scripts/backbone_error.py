"""Diagnostic (GPU): error of the image encoder (ResNet-50 + U-Net) variants against an fp64 CPU evaluation, and what
it does to the selection / outputs of the dexycb eval fixture scenario.
    python scripts/backbone_error.py"""
import copy
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from hoisdf_b200 import ops, synthetic as syn  # noqa: E402
from hoisdf_b200.config import cfg  # noqa: E402
from hoisdf_b200.model import get_model  # noqa: E402
from oracle import hoisdf_oracle as O  # noqa: E402

torch.backends.cudnn.allow_tf32 = False
torch.backends.cuda.matmul.allow_tf32 = False
dev = torch.device("cuda:0")


def rel(a, b):
    a, b = a.detach().cpu().double(), b.detach().cpu().double()
    return float((a - b).abs().max() / max(b.abs().max(), 1e-30))


def cpu_pyramid(sd, img, dtype):
    from hoisdf_b200.nets.module import BackboneNet, DecoderNet, DecoderNet_big
    bb, dec = BackboneNet(50), (DecoderNet_big() if cfg.use_big_decoder else DecoderNet())
    bb.load_state_dict({k[len("backbone_net."):]: v for k, v in sd.items() if k.startswith("backbone_net.")})
    dec.load_state_dict({k[len("decoder_net."):]: v for k, v in sd.items() if k.startswith("decoder_net.")})
    bb, dec = bb.to(dtype).eval(), dec.to(dtype).eval()
    f, s = bb(img.to(dtype))
    return dec(f, s)[0]


def variants():
    yield "cudnn+cudnn", dict(tc_backbone=False, tc_unet=False), None
    yield "cudnn+h3", dict(tc_backbone=False, tc_unet=True), None
    yield "h3+h3", dict(tc_backbone=True, tc_unet=True), None


def main():
    g = np.load(os.path.join(os.path.dirname(__file__), "..", "tests", "golden", "dexycb_eval_seed14.npz"))
    seed, B, ph, po = int(g["seed"]), int(g["batch"]), int(g["num_samp_hand"]), int(g["num_samp_obj"])
    cfg.set_setting("dexycb")
    type(cfg).dataset, type(cfg).num_samp_hand, type(cfg).num_samp_obj = "dexycb", ph, po
    sd = syn.full_state_dict(seed, "dexycb")
    model = get_model("test", mano_buffers=syn.mano_buffers(seed))
    model.load_state_dict(sd, strict=True)
    img, meta = syn.image_batch(seed, B), syn.camera_meta(seed, B)
    inputs, targets = syn.dexycb_extras(seed, B, ph, po)
    # fp64 pyramid on the CPU
    with torch.no_grad():
        pyr64 = cpu_pyramid(sd, img, torch.float64)
        otaps = {}
        oout = O.model_eval_dexycb(sd, img, inputs, targets, meta,
                                   O.default_cfg(dataset="dexycb", num_samp_hand=ph, num_samp_obj=po))
    model = model.to(dev).eval()
    to_dev = lambda d: {k: v.to(dev) for k, v in d.items()}  # noqa: E731
    orig_lin, orig_conv = ops.linear_h3, ops.conv_h3
    for name, flags, two in variants():
        for k, v in flags.items():
            setattr(type(cfg), k, v)
        with torch.no_grad():
            pyr, _ = model.run_image_encoder(img.to(dev))
            errs = {k: rel(pyr[k], pyr64[k]) for k in pyr64}
        ops.linear_h3, ops.conv_h3 = orig_lin, orig_conv
        out = model({"img": img.to(dev), **to_dev(inputs)}, to_dev(targets), to_dev(meta), "eval")
        outs = {k: rel(out[k], oout[k]) for k in ("mano_mesh_out", "mano_joints_out", "hand_joints_out")}
        print("%-14s pyramid err vs fp64: %s" % (name, " ".join("%s=%.1e" % (k[6:], v) for k, v in errs.items())))
        print("%-14s outputs vs oracle : %s" % ("", " ".join("%s=%.1e" % kv for kv in outs.items())), flush=True)
    # the oracle's own fp32 pyramid against fp64 (MKL-DNN)
    with torch.no_grad():
        p32 = cpu_pyramid(sd, img, torch.float32)
    print("cpu fp32       pyramid err vs fp64: %s" % " ".join("%s=%.1e" % (k[6:], rel(p32[k], pyr64[k])) for k in pyr64))


if __name__ == "__main__":
    main()
