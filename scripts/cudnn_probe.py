import sys, os, torch, warnings
warnings.filterwarnings("ignore")
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from hoisdf_b200 import synthetic as syn
from hoisdf_b200.config import cfg
from hoisdf_b200.nets.module import BackboneNet, DecoderNet_big
dev = torch.device("cuda:0")
sd = syn.full_state_dict(0, "ho3d")
def build(cl):
    bb, dn = BackboneNet(), DecoderNet_big()
    bb.load_state_dict({k[len("backbone_net."):]: v for k, v in sd.items() if k.startswith("backbone_net.")})
    dn.load_state_dict({k[len("decoder_net."):]: v for k, v in sd.items() if k.startswith("decoder_net.")})
    bb, dn = bb.to(dev).eval(), dn.to(dev).eval()
    if cl: bb, dn = bb.to(memory_format=torch.channels_last), dn.to(memory_format=torch.channels_last)
    return bb, dn
img = syn.image_batch(100, 32).to(dev)
for cl in (True, False):
    for tf32 in (False, True):
        for bench in (True, False):
            torch.backends.cudnn.allow_tf32 = tf32; torch.backends.cudnn.benchmark = bench
            bb, dn = build(cl)
            x = img.contiguous(memory_format=torch.channels_last) if cl else img
            with torch.no_grad():
                for _ in range(3): f, s = bb(x); p, o = dn(f, s)
                e = [torch.cuda.Event(enable_timing=True) for _ in range(3)]
                e[0].record()
                for _ in range(3): f, s = bb(x)
                e[1].record()
                for _ in range(3): p, o = dn(f, s)
                e[2].record(); torch.cuda.synchronize()
            print("channels_last=%s tf32=%s benchmark=%s  backbone %.2f ms  unet %.2f ms" % (cl, tf32, bench, e[0].elapsed_time(e[1])/3, e[1].elapsed_time(e[2])/3), flush=True)
