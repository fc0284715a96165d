import sys; sys.path.insert(0, '/root/repo')
import torch
from hoisdf_b200 import autograd as A
cuda = torch.device("cuda:0")
def rnd(seed, *shape, lo=-1.0, hi=1.0):
    g = torch.Generator().manual_seed(seed); return torch.rand(*shape, generator=g) * (hi - lo) + lo
for (m, k, n) in [(576, 256, 20), (576, 20, 256), (576, 17, 256), (576, 24, 256), (576, 32, 256), (576, 40, 256), (100, 256, 20), (576, 256, 17), (576,256,32), (64, 256, 256), (96,256,256), (34, 256, 256), (576, 60, 256)]:
    a, b = rnd(1, m, k), rnd(2, n, k, lo=-0.1, hi=0.1)
    ref = a.double() @ b.double().t()
    y = A.matmul_nt(a.to(cuda), b.to(cuda))
    err = float((y.cpu().double() - ref).abs().max() / ref.abs().max())
    print(m, k, n, "err %.3e" % err)
