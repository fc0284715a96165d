"""TMEM accumulation chunk of the ResNet-50 encoder GEMMs (cfg.backbone_chunk_kb) and of the pyramid projection
(cfg.projection_chunk_kb): time of the image encoder at the bench shape and pyramid error against an fp64 evaluation.
    python scripts/chunk_sweep.py"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from hoisdf_b200 import ops, synthetic as syn
from hoisdf_b200.config import cfg
from hoisdf_b200.model import get_model, PyramidContext
dev = torch.device("cuda:0")
cfg.set_setting("ho3d")
sd = syn.full_state_dict(0, "ho3d")
model = get_model("test", mano_buffers=syn.mano_buffers(0)); model.load_state_dict(sd, strict=True); model = model.to(dev).eval()
img = syn.image_batch(100, 32).to(dev)
# fp64 reference of 2 samples on the CPU (cuDNN-free)
from hoisdf_b200.nets.module import BackboneNet, DecoderNet_big
bb, dec = BackboneNet(50), DecoderNet_big()
bb.load_state_dict({k[len("backbone_net."):]: v for k, v in sd.items() if k.startswith("backbone_net.")})
dec.load_state_dict({k[len("decoder_net."):]: v for k, v in sd.items() if k.startswith("decoder_net.")})
bb, dec = bb.double().eval(), dec.double().eval()
with torch.no_grad():
    f, s = bb(img[:2].cpu().double()); ref = dec(f, s)[0]
print("| backbone_chunk_kb | projection_chunk_kb | encoder ms (B=32) | projection ms | worst pyramid error vs fp64 (of the level's max) |")
print("|---|---|---|---|---|")
for ck, pk in ((1, 1), (2, 1), (4, 1), (2, 2), (4, 4)):
    type(cfg).backbone_chunk_kb, type(cfg).projection_chunk_kb = ck, pk
    model._resnet_h3 = None
    with torch.no_grad():
        for _ in range(3):
            pyr, _ = model.run_image_encoder(img)
        torch.cuda.synchronize()
        e = [torch.cuda.Event(enable_timing=True) for _ in range(3)]
        e[0].record()
        for _ in range(10):
            pyr, _ = model.run_image_encoder(img)
        e[1].record()
        for _ in range(10):
            PyramidContext(pyr, model).gmaps
        e[2].record(); torch.cuda.synchronize()
        p2, _ = model.run_image_encoder(img[:2])
        err = max(float((p2[k].cpu().double() - ref[k]).abs().max() / ref[k].abs().max()) for k in ref)
    print("| %d | %d | %.3f | %.3f | %.2e |" % (ck, pk, e[0].elapsed_time(e[1]) / 10, e[1].elapsed_time(e[2]) / 10, err))
