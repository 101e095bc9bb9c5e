"""tuning aid: fused gather+score forward of the distance models with the row register cache on / off
(KGE_SCORE_CH), tables >> L2, random ids; entity-row bytes only in the algorithmic figure."""
import json, os, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, ROOT)
from pykg2vec_b200 import _lib
dev = torch.device("cuda", 0)
gen = torch.Generator(device=dev).manual_seed(0)
def run(name, N, d, n, ntab_e, extra_r):
    R = 1000
    tabs = {"transe": ["e", "r"], "transh": ["e", "r", "r"], "transd": ["e", "r", "e", "r"]}[name]
    T = [(torch.rand((N if k == "e" else R, d), device=dev, generator=gen) - 0.5) * 0.2 for k in tabs]
    desc = _lib.ModelDesc(name, T, d, l1_flag=False)
    h = torch.randint(0, N, (n,), device=dev, generator=gen); r = torch.randint(0, R, (n,), device=dev, generator=gen)
    t = torch.randint(0, N, (n,), device=dev, generator=gen); o = torch.empty(n, dtype=torch.float32, device=dev)
    for ch in ("default", "0", "4", "8"):
        if ch == "default": os.environ.pop("KGE_SCORE_CH", None)
        else: os.environ["KGE_SCORE_CH"] = ch
        for _ in range(3): _lib.score_fwd(desc, h, r, t, out=o)
        torch.cuda.synchronize()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        for _ in range(10): _lib.score_fwd(desc, h, r, t, out=o)
        b.record(); torch.cuda.synchronize()
        ms = a.elapsed_time(b) / 10
        alg = n * (2 * ntab_e * d * 4 + 28)
        print(json.dumps({"model": name, "d": d, "ch": ch, "ms": ms, "GBps_entity_rows": alg / ms / 1e6, "frac_of_6570": alg / ms / 1e6 / 6570}))
    os.environ.pop("KGE_SCORE_CH", None)
run("transe", 2_000_000, 200, 4_000_000, 1, 0)
run("transe", 4_000_000, 50, 8_000_000, 1, 0)
run("transh", 2_000_000, 200, 2_000_000, 1, 0)
run("transd", 1_000_000, 200, 2_000_000, 2, 0)
