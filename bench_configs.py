#!/usr/bin/env python
"""Per-configuration measurements for BASELINE.json configs[2..4] (NOT the driver's bench.py,
which measures configs[1]): for each model/shape, device-resident timings of
  * one fused training step through pykg2vec_b200.trainer.Trainer.train_batch internals, and
  * one 1-vs-all evaluation batch (kge_rank_1vsall, both directions, raw + filtered),
reported as scored triples/s.  One JSON line per configuration.

    python bench_configs.py [--reps 10] [--out gpurun_out/configs_r1.jsonl] [--only NAME]
"""
import argparse
import json
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
import pykg2vec_b200  # noqa: E402
from pykg2vec_b200 import _lib  # noqa: E402
from pykg2vec_b200.synthetic import SHAPES, SyntheticConfig, SyntheticKnowledgeGraph  # noqa: E402
from pykg2vec_b200.trainer import Trainer  # noqa: E402

CONFIGS = [
    # name, model, dataset shape, ctor/config kwargs, batch, neg_rate, optimizer, eval queries
    ("cfg1_transe_umls_d50", "transe", "umls", dict(hidden_size=50, l1_flag=True, margin=0.8), 128, 1, "sgd", 512),
    ("cfg2_transe_fb15k237_d200", "transe", "fb15k_237", dict(hidden_size=200, l1_flag=False, margin=5.0), 512, 1, "sgd", 512),
    ("cfg3_distmult_wn18rr_d200", "distmult", "wn18rr", dict(hidden_size=200, lmbda=1e-4), 512, 1, "adagrad", 512),
    ("cfg3_complex_wn18rr_d200", "complex", "wn18rr", dict(hidden_size=200, lmbda=1e-4), 512, 1, "adagrad", 512),
    ("cfg4_rotate_fb15k_d1000_neg256", "rotate", "fb15k", dict(hidden_size=1000, margin=24.0, alpha=1.0), 1024, 256, "adagrad", 512),
    ("cfg5_complex_yago310_d500", "complex", "yago3_10", dict(hidden_size=500, lmbda=1e-4), 512, 1, "adagrad", 512),
    # the CLI's default optimizer (-opt adam, common.py:50): fused dense Adam (kge_optim_apply_dense)
    ("cfg2_transe_fb15k237_d200_adam", "transe", "fb15k_237", dict(hidden_size=200, l1_flag=False, margin=5.0), 512, 1, "adam", 512),
    ("cfg3_complex_wn18rr_d200_adam", "complex", "wn18rr", dict(hidden_size=200, lmbda=1e-4), 512, 1, "adam", 512),
]


def time_ms(fn, reps, flush=None):
    for _ in range(2):
        fn()
    torch.cuda.synchronize()
    tot = 0.0
    for _ in range(reps):
        if flush is not None:
            flush.zero_()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        fn()
        b.record()
        torch.cuda.synchronize()
        tot += a.elapsed_time(b)
    return tot / reps


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--reps", type=int, default=10)
    ap.add_argument("--out", default=None)
    ap.add_argument("--only", default=None)
    args = ap.parse_args()
    dev = torch.device("cuda", 0)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
    lines = []
    for name, model, shape, kw, B, neg, opt, Q in CONFIGS:
        if args.only and args.only not in name:
            continue
        n_ent, n_rel = SHAPES[shape][:2]
        kg = SyntheticKnowledgeGraph(n_ent, n_rel, 4096, 64, max(Q, 64), seed=0, name=shape + "-shaped")
        cfg = SyntheticConfig(kg, device=dev, optimizer=opt, learning_rate=0.01, batch_size=B, neg_rate=neg, **kw)
        torch.manual_seed(0)
        m = pykg2vec_b200.import_model(model)(**cfg.__dict__)
        tr = Trainer(m, cfg)
        tr.build_model()
        rng = np.random.RandomState(1)
        to = lambda a: torch.from_numpy(np.ascontiguousarray(a, dtype=np.int64)).to(dev)
        if m.training_strategy.name == "PAIRWISE_BASED":
            ids = [to(rng.randint(n_ent, size=B)), to(rng.randint(n_rel, size=B)), to(rng.randint(n_ent, size=B)),
                   to(rng.randint(n_ent, size=B * neg)), to(rng.randint(n_rel, size=B * neg)),
                   to(rng.randint(n_ent, size=B * neg))]
            step = lambda: tr._fused_pairwise(ids)
            scored = B * (1 + neg)
        else:
            n = B * (1 + neg)
            ids = [to(rng.randint(n_ent, size=n)), to(rng.randint(n_rel, size=n)), to(rng.randint(n_ent, size=n)),
                   to(np.where(np.arange(n) % (1 + neg) == 0, 1, -1))]
            step = lambda: tr._fused_pointwise(ids)
            scored = n
        with torch.no_grad():
            train_ms = time_ms(step, args.reps, flush)
        desc = m.kge_desc()
        test = kg.arrays["test"][:Q]
        qh, qr, qt = to(test[:, 0]), to(test[:, 1]), to(test[:, 2])
        counts = torch.zeros((Q, 4), dtype=torch.int32, device=dev)
        ws = torch.empty(max(_lib.rank_workspace_bytes(desc, Q), 16), dtype=torch.uint8, device=dev)

        def rank():
            counts.zero_()
            _lib.rank_1vsall(desc, qh, qr, qt, None, None, counts=counts, workspace=ws)
        eval_ms = time_ms(rank, args.reps, flush)
        ntab = sum(1 for w in m.kge_tables() if w.shape[0] == n_ent)
        alg = 2 * Q * n_ent * ntab * m.kge_spec().dim * 4
        line = {"config": name, "model": model, "N": n_ent, "R": n_rel, "d": m.kge_spec().dim, "B": B, "neg": neg,
                "optimizer": opt, "train_step_ms": train_ms, "train_scored_triples_per_s": scored / train_ms * 1e3,
                "eval_Q": Q, "eval_batch_ms": eval_ms, "eval_scored_triples_per_s": 2 * Q * n_ent / eval_ms * 1e3,
                "eval_algorithmic_GBps": alg / eval_ms / 1e6,
                "table_MB": sum(w.numel() * 4 for w in m.kge_tables()) / 1e6}
        print(json.dumps(line))
        lines.append(line)
        del tr, m, ws
        torch.cuda.empty_cache()
    if args.out:
        with open(args.out, "w") as f:
            for l in lines:
                f.write(json.dumps(l) + "\n")


if __name__ == "__main__":
    main()
