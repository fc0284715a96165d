"""Developer check: U-Net decoder on the FP16x3 kernels (nets/unet_h3.py) vs the cuDNN fp32 modules."""
import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from hoisdf_b200 import ops, synthetic as syn
from hoisdf_b200.nets.module import Decoder, Decoder_big
from hoisdf_b200.nets.unet_h3 import UNetH3
torch.backends.cudnn.allow_tf32 = False
torch.backends.cuda.matmul.allow_tf32 = False
dev = torch.device("cuda:0")
B = int(sys.argv[1]) if len(sys.argv) > 1 else 4
for arch, cls in (("dexycb", Decoder), ("ho3d", Decoder_big)):
    torch.manual_seed(0)
    dec = cls()
    sd = {k[len("decoder_net.resnet_decoder."):]: v for k, v in syn.full_state_dict(3, arch).items()
          if k.startswith("decoder_net.resnet_decoder.")}
    dec.load_state_dict(sd, strict=True)
    dec = dec.to(dev).eval()
    feat = torch.relu(torch.randn(B, 2048, 8, 8, device=dev))
    skips = {"stride16": torch.relu(torch.randn(B, 1024, 16, 16, device=dev)),
             "stride8": torch.relu(torch.randn(B, 512, 32, 32, device=dev)),
             "stride4": torch.relu(torch.randn(B, 256, 64, 64, device=dev)),
             "stride2": torch.relu(torch.randn(B, 64, 128, 128, device=dev))}
    run = UNetH3(dec)
    with torch.no_grad():
        ref_pyr, ref_out = dec(feat, skips)
        pyr, out = run(feat, skips)
        torch.cuda.synchronize()
        for k in ref_pyr:
            a, r = pyr[k].double(), ref_pyr[k].double()
            print(arch, k, tuple(pyr[k].shape), "rel.err %.2e" % float((a - r).abs().max() / r.abs().max()), flush=True)
        print(arch, "decoder_out rel.err %.2e" % float((out.double() - ref_out.double()).abs().max() / ref_out.abs().max()))
        for fn, tag in ((lambda: dec(feat, skips), "cuDNN fp32"), (lambda: run(feat, skips), "FP16x3")):
            for _ in range(2): fn()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(3): fn()
            e1.record(); torch.cuda.synchronize()
            print(arch, tag, "%.3f ms / forward (B=%d)" % (e0.elapsed_time(e1) / 3, B), flush=True)
