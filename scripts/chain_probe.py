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

if os.environ.get("HOISDF_CHAIN_PROF"):
    import ctypes as C
    from hoisdf_b200 import _capi
    buf = (C.c_longlong * 16)()
    _capi.lib.hoisdf_debug_chain_profile(buf)            # reset
    fn()
    torch.cuda.synchronize()
    _capi.lib.hoisdf_debug_chain_profile(buf)
    v = list(buf)
    tiles = (rows + 127) // 128 / 148.0
    print("CTA 0, one launch (%.1f tiles): MMA issuer total %d cycles = %.0f / tile; waits: weights %.1f%%, accumulator %.1f%%, "
          "A operand %.1f%%" % (tiles, v[3], v[3] / tiles, 100.0 * v[0] / v[3], 100.0 * v[1] / v[3], 100.0 * v[2] / v[3]))
    print("epilogue warp: total %d, waiting for accumulators %.1f%%;  weight producer: total %d, waiting for a free stage %.1f%%"
          % (v[5], 100.0 * v[4] / max(v[5], 1), v[7], 100.0 * v[6] / max(v[7], 1)))
    stages = {0: 16, 1: 20, 2: 16, 3: 36, 4: 32}
    mmas = {0: 64, 1: 76, 2: 64, 3: 136, 4: 128}
    for l, name in enumerate(("s1 (SS)", "linh0 (SS)", "linh1 (TS)", "linh2 (SS)", "linh3 (TS)")):
        per_tile = v[8 + l] / tiles
        print("  %-11s issue path %.0f cycles / tile = %.0f per stage, %.0f per MMA" % (name, per_tile, per_tile / stages[l], per_tile / mmas[l]))
