"""BASELINE config 5: SDF-MLP (SDFDecoder.forward) in isolation, point-count sweep 256/1024/4096/16384 at batch 32.
Rows = 32 * points, input (rows, 289) ~ N(0,1).  Reports time, rows/s, algorithmic TFLOP/s (1 573 888 FLOP/row) and
algorithmic HBM GB/s (1 160 B/row: 289 floats in, 1 out) against the measured peaks, for the tcgen05 FP16x3 path (default),
the tcgen05 3xTF32 path and the fp32 FMA path.  Developer tool: prints a markdown table (committed under profiles/)."""
import json, os, sys, torch, warnings
warnings.filterwarnings("ignore")
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from hoisdf_b200 import ops, synthetic as syn
from hoisdf_b200.nets.sdf_net import SDFDecoder
from oracle import hoisdf_oracle as O
dev = torch.device("cuda:0")
peaks = json.load(open(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "MEASURED_PEAKS.json"))) \
    if os.path.exists("MEASURED_PEAKS.json") else {"hbm_gbs": 6650.0, "bf16_tflops": 1590.0}
sd = syn.hot_path_state_dict(0, "ho3d")
dec = SDFDecoder(256, 33).to(dev).eval()
dec.load_state_dict({k[len("hand_sdf_decoder."):]: v for k, v in sd.items() if k.startswith("hand_sdf_decoder.")})
FLOP, BYTES = ops.SDF_DECODER_FLOPS, 1160.0
print("| points/sample | rows | impl | ms | M rows/s | TFLOP/s (algorithmic) | % of bf16 burst peak | GB/s (algorithmic) | % of HBM peak | max abs err vs oracle |")
print("|---|---|---|---|---|---|---|---|---|---|")
for pts in ((int(sys.argv[1]),) if len(sys.argv) > 1 else (256, 1024, 4096, 16384)):
    rows = 32 * pts
    g = torch.Generator().manual_seed(pts)
    x = torch.randn(rows, 289, generator=g)
    xd = x.to(dev)
    ref = O.sdf_decoder(sd, "hand_sdf_decoder", x[:4096])
    for impl in ("fused tcgen05 kernel, 1 fp16 product (screening)", "tcgen05 FP16x3", "tcgen05 3xTF32", "fp32 FMA"):
        fused = impl.startswith("fused")
        ops.USE_TENSOR_CORES = impl.startswith("tc") or fused
        ops.TC_MODE = "tf32" if "TF32" in impl else "h3"
        dec._packed = None
        if fused:
            # the ONE persistent kernel (csrc/sdf_chain.cu, decoder mode): input rows already in the fp16 row format, like
            # inside sdf_infer where the previous kernel's epilogue writes them; 4 bytes per row come back
            buf = ops.sdf_pad_input(xd)
            o1 = torch.empty(rows, device=dev)
            run = lambda: (ops.sdf_chain(dec.packed(), o1, x=buf).view(-1, 1), None)
        else:
            run = lambda: dec(xd)
        with torch.no_grad():
            for _ in range(3): out, _ = run()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(10): out, _ = run()
            e1.record(); torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / 10
        err = float((out[:4096].cpu() - ref).abs().max())
        tf = rows * FLOP / ms / 1e9; gbs = rows * BYTES / ms / 1e6
        print("| %d | %d | %s | %.3f | %.1f | %.1f | %.1f%% | %.1f | %.2f%% | %.1e |" % (
            pts, rows, impl, ms, rows / ms / 1e3, tf, 100 * tf / peaks["bf16_tflops"], gbs, 100 * gbs / peaks["hbm_gbs"], err))
