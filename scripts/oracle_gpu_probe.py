import sys, time, torch
sys.path.insert(0, "/root/repo")
from hoisdf_b200 import synthetic as syn
from oracle import hoisdf_oracle as O
torch.backends.cuda.matmul.allow_tf32 = False; torch.backends.cudnn.allow_tf32 = False
dev = torch.device("cuda:0")
sd = {k: v.to(dev) for k, v in syn.full_state_dict(0, "ho3d").items()}
ocfg = O.default_cfg(num_samp_hand=1536, num_samp_obj=512)
B = 4
img = syn.image_batch(1000, B).to(dev); meta = {k: v.to(dev) for k, v in syn.camera_meta(1000, B).items()}
with torch.no_grad():
    for i in range(3):
        torch.cuda.synchronize(); t0 = time.perf_counter()
        out = O.model_eval(sd, img, meta, ocfg, "ho3d")
        torch.cuda.synchronize(); print("B=%d %.3f s" % (B, time.perf_counter() - t0))
print({k: tuple(v.shape) for k, v in out.items()})
