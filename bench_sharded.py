#!/usr/bin/env python
"""Multi-GPU evaluation measurements for BASELINE.json configs[3] and configs[4] (NOT the
driver's bench.py).  Launch with torchrun, one rank per GPU:

  python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 \
      --master-port 29533 bench_sharded.py [--queries 5000]

* config 5 — ComplEx, YAGO3-10 shape (N=123,182, R=37, d=500): the two entity tables are
  ROW-PARTITIONED across the ranks (pykg2vec_b200.sharding.RowShardedRanker): an all-gather of the
  query rows (each contributed by its owner) assembles the compact query table, every rank sweeps its
  own rows for all queries with kge_rank_1vsall (tensor-core two-level sweep), ONE all-gather of the
  [Q,4] partial rank counts, summed locally.  Checked against the unsharded kernel result on rank 0
  (exact).  Also reachable as `bench.py --config 5` (and `--config 4`).
* config 4 — RotatE, FB15k shape (N=14,951, R=1,345, d=1000): tables replicated, the test triples
  are sharded, one all-gather of the ranks.
Times are device-side (CUDA events), max over ranks; rank 0 prints one JSON line per config.
"""
import argparse
import json
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
from pykg2vec_b200 import _lib, sharding  # noqa: E402


def timed(fn, reps):
    fn()
    torch.cuda.synchronize()
    best = []
    for _ in range(reps):
        if dist.is_initialized():
            dist.barrier()
        torch.cuda.synchronize()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        out = fn()
        b.record()
        torch.cuda.synchronize()
        best.append(a.elapsed_time(b))
    ms = float(np.median(best))
    if dist.is_initialized():
        t = torch.tensor([ms], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms = float(t.item())
    return ms, out


def main(argv=None, only=None):
    ap = argparse.ArgumentParser()
    ap.add_argument("--queries", type=int, default=5000)
    ap.add_argument("--reps", type=int, default=3)
    args = ap.parse_args(argv)
    rank, world = sharding.init_distributed()
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    Q = args.queries
    lines = []
    rng = np.random.RandomState(1)

    # ---------------- config 5: ComplEx YAGO3-10 shape, row-sharded entity tables ----------------
    if only in (None, 5):
        N, R, d = 123182, 37, 500
        g = torch.Generator(device="cpu").manual_seed(0)
        bound = (6.0 / (N + d)) ** 0.5
        full = [(torch.rand((N, d), generator=g) * 2 - 1) * bound, (torch.rand((N, d), generator=g) * 2 - 1) * bound]
        rel = [((torch.rand((R, d), generator=g) * 2 - 1) * 0.4).to(dev), ((torch.rand((R, d), generator=g) * 2 - 1) * 0.4).to(dev)]
        qh, qr, qt = rng.randint(N, size=Q), rng.randint(R, size=Q), rng.randint(N, size=Q)
        lo, hi = sharding.shard_range(N, world, rank)
        ent_local = [t[lo:hi].contiguous().to(dev) for t in full]
        ranker = sharding.RowShardedRanker(sharding.cuda_count_fn("complex", d), N, ent_local, rel, (0, 1), (2, 3),
                                           rank=rank, world=world)
        ms, counts = timed(lambda: ranker.rank_queries(qh, qr, qt), args.reps)
        ok = None
        if rank == 0:
            desc = _lib.ModelDesc("complex", [t.to(dev) for t in full] + rel, d)
            to = lambda a: torch.from_numpy(np.ascontiguousarray(a)).to(dev)
            want = _lib.rank_1vsall(desc, to(qh), to(qr), to(qt))
            ok = bool(torch.equal(want, counts))
            del desc
        lines.append({"config": "cfg5_complex_yago310_d500_row_sharded", "n_gpus": world, "Q": Q, "N": N, "d": d,
                      "rows_per_gpu": hi - lo, "shard_MB_per_gpu": 2 * (hi - lo) * d * 4 / 1e6, "ms": ms,
                      "scored_triples_per_s": 2.0 * Q * N / ms * 1e3, "matches_unsharded_kernel": ok,
                      "collectives": "2 all-gather (owners' query rows, %d x %d fp32, one per entity table) + 1 all-gather ([Q,4] int32 partial counts)"
                                     % (len(np.unique(np.concatenate([qh, qt]))), d)})
        del full, ent_local, ranker
        torch.cuda.empty_cache()

    # ---------------- config 4: RotatE FB15k shape, replicated tables, sharded test triples -------
    if only in (None, 4):
        N, R, d, margin = 14951, 1345, 1000, 24.0
        er = (margin + 2.0) / d
        g = torch.Generator(device="cpu").manual_seed(2)
        tabs = [((torch.rand((N, d), generator=g) * 2 - 1) * er).to(dev), ((torch.rand((N, d), generator=g) * 2 - 1) * er).to(dev),
                ((torch.rand((R, d), generator=g) * 2 - 1) * er).to(dev)]
        desc = _lib.ModelDesc("rotate", tabs, d, margin=margin, phase_scale=float(np.float32(np.pi / er)))
        qh, qr, qt = rng.randint(N, size=Q), rng.randint(R, size=Q), rng.randint(N, size=Q)
        qlo, qhi = sharding.shard_range(Q, world, rank)
        to = lambda a: torch.from_numpy(np.ascontiguousarray(a)).to(dev)
        lh, lr, lt = to(qh[qlo:qhi]), to(qr[qlo:qhi]), to(qt[qlo:qhi])

        def run():
            local_counts = _lib.rank_1vsall(desc, lh, lr, lt)
            return sharding.gather_query_shards(local_counts, Q)
        ms, allc = timed(run, args.reps)
        lines.append({"config": "cfg4_rotate_fb15k_d1000_query_sharded", "n_gpus": world, "Q": Q, "N": N, "d": d, "ms": ms,
                      "scored_triples_per_s": 2.0 * Q * N / ms * 1e3, "gathered_rows": int(allc.shape[0]),
                      "collectives": "1 all-gather ([Q,4] int32)"})
    if rank == 0:
        for l in lines:
            print(json.dumps(l), flush=True)
    if dist.is_initialized():
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
