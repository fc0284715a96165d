"""Time the fused candidate chain kernel (csrc/sdf_chain.cu) against the unfused single-product launches it replaces.
Developer tool: prints a markdown table (committed under profiles/)."""
import json, os, sys, torch, warnings
warnings.filterwarnings("ignore")
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from hoisdf_b200 import ops, synthetic as syn
from hoisdf_b200.nets.sdf_net import SDFDecoder
dev = torch.device("cuda:0")
pk = os.path.join(ROOT, "MEASURED_PEAKS.json")
peaks = json.load(open(pk)) if os.path.exists(pk) else {"hbm_gbs": 6650.0, "bf16_tflops": 1590.0}
sd = syn.hot_path_state_dict(0, "ho3d")
dec = SDFDecoder(256, 33).to(dev).eval()
dec.load_state_dict({k[len("hand_sdf_decoder."):]: v for k, v in sd.items() if k.startswith("hand_sdf_decoder.")})
pw = ops.PackedLinear.pack(sd["linear_sdfin.layers.1.weight"].to(dev), sd["linear_sdfin.layers.1.bias"].to(dev))
packed = dec.packed()


def timed(fn, n=10):
    for _ in range(3):
        fn()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n


print("| rows | what | ms | M rows/s | TFLOP/s executed | % of bf16 burst peak |")
print("|---|---|---|---|---|---|")
for rows in (8192, 32768, 131072, 524288, 628248):
    g = torch.Generator().manual_seed(rows)
    h = torch.rand(rows, 512, generator=g).mul_(0.5).to(dev)
    idx = torch.randint(0, 64 ** 3, (rows,), generator=g, dtype=torch.int32).to(dev)
    hs = ops.split_rows(h)
    rs = ops.SplitRows.empty(rows, ops.ROWH_LD, dev)
    ha, hb = ops.SplitRows.empty(rows, 512, dev), ops.SplitRows.empty(rows, 512, dev)
    out = torch.empty(rows, device=dev)
    f_dec = ops.SDF_DECODER_FLOPS
    f_all = f_dec + 2.0 * 512 * 256

    def unfused_all():
        ops.linear(hs, pw, ops.ACT_RELU, out=rs.window(0, 256), chunk_kb=ops.SCREEN_CHUNK_KB, single=True)
        ops.posenc(rs, lattice_index=idx, bins=64)
        ops.sdf_decoder(packed, rs, h_a=ha, h_b=hb, out=out, chunk_kb=ops.SCREEN_CHUNK_KB, single=True)

    unfused_all()
    cases = [
        ("fused chain, rows mode (sdfin.1 + posenc + decoder)", lambda: ops.sdf_chain(packed, out, sdfin1=pw, a0=hs, lattice_index=idx), f_all),
        ("unfused single-product launches (same work)", unfused_all, f_all),
        ("fused chain, decoder mode (SDFDecoder only)", lambda: ops.sdf_chain(packed, out, x=rs), f_dec),
        ("unfused single-product SDFDecoder", lambda: ops.sdf_decoder(packed, rs, h_a=ha, h_b=hb, out=out, chunk_kb=ops.SCREEN_CHUNK_KB, single=True), f_dec),
        ("FP16x3 SDFDecoder (fp32-grade, 3 products)", lambda: ops.sdf_decoder(packed, rs, h_a=ha, h_b=hb, out=out), f_dec),
    ]
    for name, fn, fl in cases:
        ms = timed(fn)
        tf = rows * fl / ms / 1e9
        print("| %d | %s | %.3f | %.1f | %.1f | %.1f%% |" % (rows, name, ms, rows / ms / 1e3, tf, 100 * tf / peaks["bf16_tflops"]))
