"""Stand-alone launches of the FP16x3 GEMM at the shapes the bench step uses (for targeted ncu captures and timing).
    python scripts/h3_shapes.py [--time]      # --time: CUDA-event timing table instead of the profiler bracket"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from hoisdf_b200 import ops  # noqa: E402
from hoisdf_b200.nets.unet_h3 import TAPS_3X3  # noqa: E402

dev = torch.device("cuda:0")
g = torch.Generator(device="cpu").manual_seed(0)


def rnd(*shape, s=1.0):
    return (torch.rand(*shape, generator=g) * 2 - 1).mul_(s).to(dev)


def lin(m, n, k, split=True, residual=False, chunk=0, act=1, single=False):
    x = ops.split_rows(rnd(m, k))
    pw = ops.PackedLinearH3.pack(rnd(n, k, s=0.1), rnd(n), chunk_kb=chunk)
    res = ops.split_rows(rnd(m, n)) if residual else None
    out = ops.SplitRows.empty(m, n, dev) if split else torch.empty(m, n, device=dev)
    name = "lin M=%d N=%d K=%d %s%s chunk=%d%s" % (m, n, k, "split" if split else "f32", " +res" if residual else "", chunk,
                                                  " single" if single else "")
    return name, 2.0 * m * n * k, lambda: ops.linear_h3(x, pw, act, out=out, residual_split=res, single=single)


def conv(b, h, cin, cout, stride=1, chunk=0):
    x = ops.split_rows(rnd(b * h * h, cin))
    pw = ops.PackedLinearH3.pack(rnd(cout, 9 * cin, s=0.05), rnd(cout), chunk_kb=chunk)
    ho = h // stride
    out = ops.SplitRows.empty(b * ho * ho, cout, dev)
    name = "conv B=%d %dx%d Cin=%d Cout=%d s=%d chunk=%d" % (b, h, h, cin, cout, stride, chunk)
    return name, 2.0 * b * ho * ho * cout * 9 * cin, lambda: ops.conv_h3(x, b, h, h, cin, pw, TAPS_3X3, ho, ho,
                                                                          stride=stride, act=1, out=out)


CASES = [
    lin(628248, 512, 512, chunk=1 << 20, single=True),
    lin(131072, 256, 64, residual=True, chunk=1),
    lin(524288, 64, 128),
    lin(65536, 1024, 256),
    lin(65536, 256, 1024, split=False, act=0),
    lin(628248, 512, 512, chunk=1 << 20),
    lin(628248, 512, 512),
    conv(32, 64, 64, 64, chunk=1),
    conv(32, 32, 1024, 512),
]

for _, _, fn in CASES:
    fn()
torch.cuda.synchronize()
if "--time" in sys.argv:
    for name, flops, fn in CASES:
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        fn()
        e0.record()
        for _ in range(5):
            fn()
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / 5
        print("%-52s %8.3f ms %7.1f TF/s" % (name, ms, flops / ms / 1e9))
else:
    torch.cuda.profiler.start()
    for _, _, fn in CASES:
        fn()
    torch.cuda.synchronize()
    torch.cuda.profiler.stop()
