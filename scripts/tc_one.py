import sys, os, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from hoisdf_b200 import ops
dev = torch.device("cuda:0")
m, n, k = int(sys.argv[1]), int(sys.argv[2]), int(sys.argv[3])
x = torch.randn(m, k, device=dev); w = torch.randn(n, k, device=dev) * 0.05; b = torch.randn(n, device=dev)
pw = ops.PackedLinear.pack(w, b); out = torch.empty(m, n, device=dev)
for _ in range(3): ops.linear(x, pw, 1, out=out)
torch.cuda.synchronize()
