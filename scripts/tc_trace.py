import sys, os, torch, ctypes
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from hoisdf_b200 import ops, _capi
dev = torch.device("cuda:0")
m, n, k = int(sys.argv[1]), int(sys.argv[2]), int(sys.argv[3]); cta = int(sys.argv[4])
passes = int(sys.argv[5]) if len(sys.argv) > 5 else 3
x = torch.randn(m, k, device=dev); w = torch.randn(n, k, device=dev) * 0.05; b = torch.randn(n, device=dev)
pw = ops.PackedLinear.pack(w, b); out = torch.empty(m, n, device=dev)
for _ in range(3): ops.linear(x, pw, 1, out=out, passes=passes)
torch.cuda.synchronize()
buf = torch.zeros(256, dtype=torch.int64, device=dev)
lib = ctypes.CDLL(_capi.LIB_PATH)
_capi.lib.hoisdf_debug_tc_trace = _capi.lib.hoisdf_debug_tc_trace if hasattr(_capi.lib, "hoisdf_debug_tc_trace") else None
f = ctypes.CDLL(_capi.LIB_PATH).hoisdf_debug_tc_trace
f.argtypes = [ctypes.c_void_p, ctypes.c_int]
_capi.lib.hoisdf_debug_tc_trace.argtypes = [ctypes.c_void_p, ctypes.c_int]
_capi.lib.hoisdf_debug_tc_trace(buf.data_ptr(), cta)
ops.linear(x, pw, 1, out=out, passes=passes)
torch.cuda.synchronize()
_capi.lib.hoisdf_debug_tc_trace(None, 0)
t = buf.cpu().tolist(); t0 = t[0]
rel = lambda v: (v - t0) / 1000.0 if v else float("nan")
print("setup done %.2f us | acc complete %.2f | epilogue done %.2f | exit %.2f" % (rel(t[1]), rel(t[2]), rel(t[3]), rel(t[4])))
nkb = min(40, (k + 15) // 16)
print("kb : producer_free  tma_landed  split_done  mma_issued   (us since CTA start)")
for kb in range(nkb):
    print("%2d : %8.2f %10.2f %10.2f %10.2f" % (kb, rel(t[8 + kb]), rel(t[48 + kb]), rel(t[88 + kb]), rel(t[128 + kb])))
