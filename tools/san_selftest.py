import os, sys, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, ROOT)
import nerfsos_b200
from nerfsos_b200 import _lib
L = _lib.lib(); dev = "cuda:0"
N, K, tm, mode = [int(x) for x in sys.argv[1:5]]
g = torch.Generator().manual_seed(0)
a = (torch.rand(128, K, generator=g) * 4).to(dev); w = (torch.randn(N, K, generator=g) * 0.3).to(dev)
d = torch.full((128, N), float("nan"), device=dev); scratch = torch.zeros(1 << 20, dtype=torch.uint8, device=dev)
rc = L.nsos_selftest_umma(_lib.ptr(a), _lib.ptr(w), _lib.ptr(d), N, K, tm, mode, _lib.ptr(scratch), scratch.numel(), None)
torch.cuda.synchronize()
ref = a.double() @ w.double().T
print("rc", rc, "max err", (d.double() - ref).abs().max().item())
