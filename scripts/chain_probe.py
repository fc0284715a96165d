"""One shape of the fused chain kernel (rows mode, 628 248 rows), a few timed launches: used under env knobs / ncu."""
import os, sys, torch, warnings
warnings.filterwarnings("ignore")
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from hoisdf_b200 import ops, synthetic as syn
from hoisdf_b200.nets.sdf_net import SDFDecoder
dev = torch.device("cuda:0")
rows = int(sys.argv[1]) if len(sys.argv) > 1 else 628248
n = int(sys.argv[2]) if len(sys.argv) > 2 else 10
sd = syn.hot_path_state_dict(0, "ho3d")
dec = SDFDecoder(256, 33).to(dev).eval()
dec.load_state_dict({k[len("hand_sdf_decoder."):]: v for k, v in sd.items() if k.startswith("hand_sdf_decoder.")})
pw = ops.PackedLinear.pack(sd["linear_sdfin.layers.1.weight"].to(dev), sd["linear_sdfin.layers.1.bias"].to(dev))
packed = dec.packed()
g = torch.Generator().manual_seed(rows)
hs = ops.split_rows(torch.rand(rows, 512, generator=g).mul_(0.5).to(dev))
idx = torch.randint(0, 64 ** 3, (rows,), generator=g, dtype=torch.int32).to(dev)
out = torch.empty(rows, device=dev)
fn = lambda: ops.sdf_chain(packed, out, sdfin1=pw, a0=hs, lattice_index=idx)
for _ in range(3):
    fn()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(n):
    fn()
e1.record()
torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / n
fl = ops.SDF_DECODER_FLOPS + 2.0 * 512 * 256
print("CL=%s STAGES=%s rows=%d  %.3f ms  %.1f TFLOP/s" % (os.environ.get("HOISDF_CHAIN_CL", "-"), os.environ.get("HOISDF_CHAIN_STAGES", "-"), rows, ms, rows * fl / ms / 1e9))
