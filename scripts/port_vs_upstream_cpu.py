"""Cross-timing of the oracle PORT against the real upstream Model.forward on the same host cores (build container only:
needs /root/reference).  VERDICT r1 weak-8: the bench's reference arm is the port, so its speed relative to the unmodified
upstream code has to be on record.  Same seeded weights / inputs, eval forward, ho3d arch, P = 1536 + 512 (configs[1]).
    python scripts/port_vs_upstream_cpu.py [batch]"""
import os, sys, time, json, warnings
warnings.filterwarnings("ignore")
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from hoisdf_b200 import synthetic as syn
from oracle import hoisdf_oracle as O, reference_shim as rs

B = int(sys.argv[1]) if len(sys.argv) > 1 else 2
arch, seed, ph, po = "ho3d", 0, 1536, 512
torch.set_num_threads(os.cpu_count() or 1)
ns = rs.load(arch)
cfg = ns["cfg"]
type(cfg).num_samp_hand, type(cfg).num_samp_obj, type(cfg).dataset = ph, po, "ho3d"
model = rs.build_model(ns, syn.mano_buffers(seed))
sd = syn.full_state_dict(seed, arch)
model.load_state_dict(sd, strict=True)
model.eval()
img, meta = syn.image_batch(1000, B), syn.camera_meta(1000, B)
ocfg = O.default_cfg(num_samp_hand=ph, num_samp_obj=po)


def best(fn, n=2):
    fn()
    ts = []
    for _ in range(n):
        t0 = time.perf_counter(); fn(); ts.append(time.perf_counter() - t0)
    return min(ts)


with torch.no_grad():
    t_up = best(lambda: model({"img": img}, syn.eval_targets(B), meta, "eval"))
    t_port = best(lambda: O.model_eval(sd, img, meta, ocfg, arch))
print(json.dumps({"batch": B, "cores": torch.get_num_threads(), "upstream_s_per_step": t_up, "port_s_per_step": t_port,
                  "upstream_samples_per_s": B / t_up, "port_samples_per_s": B / t_port, "port_over_upstream": t_up / t_port}))
