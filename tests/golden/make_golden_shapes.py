#!/usr/bin/env python
"""Reference outputs at the BASELINE.json shapes (N = 14,541 / 40,943 / 14,951 / 123,182).

Run in the build container only (needs /root/reference):
    python tests/golden/make_golden_shapes.py [case ...]

The tables of these cases are tens to hundreds of MB, so they are NOT stored: both this script and
the tests regenerate them from a seed (tests/golden_util.py: baseline_tables).  Stored per case
(tests/golden/shapes_<case>.npz): the seed, triple ids, the scores the UNMODIFIED reference model
returns for them, and for a few queries the (trank, ftrank, hrank, fhrank) of the reference's own
Evaluator.test_*_rank + MetricCalculator walk with the filter dictionaries used.
"""
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
sys.path.insert(0, os.path.join(HERE, ".."))
import make_golden as mg   # noqa: E402  (installs the reference import stubs)
import golden_util as gu   # noqa: E402

N_TRIPLES = 512
N_QUERIES = {"cfg2_transe_fb15k237": 24, "cfg3_distmult_wn18rr": 12, "cfg3_complex_wn18rr": 12,
             "cfg4_rotate_fb15k": 8, "cfg5_complex_yago310": 6}


def reference_model(spec, tables):
    cls, keys = mg.SPECS[spec["model"]]
    kw = dict(hidden_size=spec["d"])
    if spec["model"] == "transe":
        kw["l1_flag"] = spec["l1"]
    elif spec["model"] == "rotate":
        kw["margin"] = spec["margin"]
    else:
        kw["lmbda"] = 0.1
    m = cls(tot_entity=spec["N"], tot_relation=spec["R"], **kw)
    m.load_state_dict({k + ".weight": torch.from_numpy(t) for k, t in zip(keys, tables)})
    m.eval()
    return m


def make(name):
    spec = gu.BASELINE_SHAPES[name]
    tables = gu.baseline_tables(spec)
    m = reference_model(spec, tables)
    rng = np.random.RandomState(spec["seed"] + 1)
    N, R = spec["N"], spec["R"]
    h = rng.randint(N, size=N_TRIPLES).astype(np.int64)
    r = rng.randint(R, size=N_TRIPLES).astype(np.int64)
    t = rng.randint(N, size=N_TRIPLES).astype(np.int64)
    with torch.no_grad():
        scores = m(torch.from_numpy(h), torch.from_numpy(r), torch.from_numpy(t)).numpy().copy()
    nq = N_QUERIES[name]
    q = [(int(h[i]), int(r[i]), int(t[i])) for i in range(nq)]
    hr_t, tr_h = mg.random_filters(rng, N, R, q, extra=30)
    ranks = mg.reference_ranks(m, N, q, hr_t, tr_h)
    out = {"seed": np.asarray(spec["seed"]), "h": h, "r": r, "t": t, "scores": scores, "ranks": ranks,
           "table0_head": tables[0][:2, :8].copy()}   # a few values to prove both sides built the same tables
    out["filt_t_ptr"], out["filt_t_idx"] = mg.csr(hr_t, [(a, b) for a, b, c in q])
    out["filt_h_ptr"], out["filt_h_idx"] = mg.csr(tr_h, [(c, b) for a, b, c in q])
    np.savez_compressed(gu.shape_case_path(name), **out)
    print("wrote", name, "scores[:3]", scores[:3], "ranks[:2]", ranks[:2].tolist())


if __name__ == "__main__":
    torch.set_num_threads(8)
    only = sys.argv[1:]
    for name in gu.BASELINE_SHAPES:
        if not only or name in only:
            make(name)
