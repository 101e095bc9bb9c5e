#!/usr/bin/env python
"""TransH / TransD full-test evaluation on the FB15k-237 shape (N=14,541, R=237, d=200, 20,466 test
triples): the model's own gather sweep vs the relation-grouped path (kge_project_entities + TransE's
tiled sweep per relation).  Host ids in, ranks out; the two must agree exactly.  One JSON line each.

    python bench_grouped.py [--out gpurun_out/grouped_r1.jsonl] [--queries 20466]
"""
import argparse
import json
import os
import sys
import time
import types

import numpy as np
import torch

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
from pykg2vec_b200 import _lib  # noqa: E402
from pykg2vec_b200.evaluator import Evaluator  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--out", default=None)
    ap.add_argument("--queries", type=int, default=20466)
    ap.add_argument("--models", default="transh,transd")
    args = ap.parse_args()
    N, R, d, Q = 14541, 237, 200, args.queries
    dev = torch.device("cuda", 0)
    rng = np.random.RandomState(0)
    hs, rs, ts = rng.randint(N, size=Q), rng.randint(R, size=Q), rng.randint(N, size=Q)
    lines = []
    for name in args.models.split(","):
        gen = torch.Generator(device=dev).manual_seed(1)
        shapes = [(N, d), (R, d), (R, d)] if name == "transh" else [(N, d), (R, d), (N, d), (R, d)]
        tabs = [(torch.rand(s, device=dev, generator=gen) - 0.5) * 0.2 for s in shapes]
        desc = _lib.ModelDesc(name, tabs, d, l1_flag=False)
        model = types.SimpleNamespace(model_name=name, kge_desc=lambda: desc, kge_tables=lambda: desc.tables)
        results = {}
        for grouped in (False, True):
            ev = object.__new__(Evaluator)
            ev.model = model
            ev.config = types.SimpleNamespace(device="cuda", tot_entity=N, relation_grouped_eval=grouped)
            ev._filter_cache, ev._workspace = {}, None
            ev.rank_triples(hs[:2048], rs[:2048], ts[:2048])          # warm-up (graphs, allocator)
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            results[grouped] = ev.rank_triples(hs, rs, ts)
            torch.cuda.synchronize()
            ms = (time.perf_counter() - t0) * 1e3
            line = {"model": name, "path": "relation-grouped (project + TransE tiled sweep)" if grouped else
                    "gather sweep", "queries": Q, "N": N, "R": R, "d": d, "ms": ms,
                    "scored_per_s": 2.0 * Q * N / ms * 1e3, "timing": "wall clock around Evaluator.rank_triples"}
            print(json.dumps(line))
            lines.append(line)
        assert np.array_equal(results[False], results[True]), "grouped ranks differ from the gather sweep"
    if args.out:
        with open(args.out, "w") as f:
            for l in lines:
                f.write(json.dumps(l) + "\n")


if __name__ == "__main__":
    main()
