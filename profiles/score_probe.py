"""ncu target: the fused gather+score forward kernels north_star names (TransE d=200 — the persistent
cp.async-staged kernel —, ComplEx d=200) on tables >> L2 with random ids (same shapes as bench.py's
rooflines_extra)."""
import os, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, ROOT)
from pykg2vec_b200 import _lib
dev = torch.device("cuda", 0)
gen = torch.Generator(device=dev).manual_seed(0)
for name, N, ne, nr, n in (("transe", 2_000_000, 1, 1, 4_000_000), ("complex", 1_000_000, 2, 2, 2_000_000)):
    d, R = 200, 1000
    tabs = [(torch.rand((N, d), device=dev, generator=gen) - 0.5) * 0.2 for _ in range(ne)] + \
           [(torch.rand((R, d), device=dev, generator=gen) - 0.5) * 0.2 for _ in range(nr)]
    desc = _lib.ModelDesc(name, tabs, d, l1_flag=False)
    h = torch.randint(0, N, (n,), device=dev, generator=gen); r = torch.randint(0, R, (n,), device=dev, generator=gen)
    t = torch.randint(0, N, (n,), device=dev, generator=gen); o = torch.empty(n, dtype=torch.float32, device=dev)
    for _ in range(3):
        _lib.score_fwd(desc, h, r, t, out=o)
    torch.cuda.synchronize()
    del tabs, desc
    torch.cuda.empty_cache()
