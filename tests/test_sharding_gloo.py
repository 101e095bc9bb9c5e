"""world_size-2 gloo tests (CPU) of the multi-GPU host logic in pykg2vec_b200/sharding.py.
The per-shard counting is injected: here the ORACLE stands in for the CUDA kernel (tests may
use the oracle as the checker), so what is under test is the partitioning, the query-row
exchange and the single all-reduce of partial counts."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _oracle_count_fn(name, dim, num_ent):
    import oracle

    def fn(shard_tables, query_tables, row_lo, row_hi, qh_c, qr, qt_c, tgt_h, tgt_t, filt_t, filt_h):
        # rebuild full-height tables holding ONLY the rows this rank may legally read: its shard
        # and the exchanged query rows; everything else is NaN so that an illegal read poisons
        # the comparison
        full = []
        for s, q in zip(shard_tables, query_tables):
            if s.shape[0] == row_hi - row_lo and s.shape[0] != q.shape[0] or s is not q and s.shape[0] == row_hi - row_lo:
                f = torch.full((num_ent, s.shape[1]), float("nan"), dtype=torch.float32)
                f[row_lo:row_hi] = s
                f[tgt_h] = q[qh_c]
                f[tgt_t] = q[qt_c]
                full.append(f.numpy())
            else:
                full.append(s.numpy())
        om = oracle.Model(name, full, dim)
        ft = (filt_t[0].numpy(), filt_t[1].numpy()) if filt_t is not None else None
        fh = (filt_h[0].numpy(), filt_h[1].numpy()) if filt_h is not None else None
        c = oracle.rank_1vsall(om, tgt_h.numpy(), qr.numpy(), tgt_t.numpy(), ft, fh, row_lo=row_lo, row_hi=row_hi)
        return torch.from_numpy(c)
    return fn


def _worker(rank, world, port, tmp):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    import sys
    sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
    sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
    from pykg2vec_b200 import sharding
    import gpu_util as gpu
    import oracle
    r, w = sharding.init_distributed(backend="gloo")
    assert (r, w) == (rank, world)
    # ---- partition helpers
    covered = []
    for k in range(world):
        lo, hi = sharding.shard_range(1001, world, k)
        covered.extend(range(lo, hi))
    assert covered == list(range(1001))
    # ---- replicated-table eval: query shards gathered back in order
    total = 37
    lo, hi = sharding.shard_range(total, world, rank)
    local = torch.arange(lo, hi, dtype=torch.int32).repeat_interleave(4).reshape(-1, 4)
    full = sharding.gather_query_shards(local, total)
    assert full[:, 0].tolist() == list(range(total))
    # ---- data-parallel batch ids: every rank sees the same global batch
    ids = torch.full((6, 5), rank, dtype=torch.int64)
    g = sharding.allgather_batch_ids(ids)
    assert g.shape == (6, 5 * world) and g[0].tolist() == [0] * 5 + [1] * 5
    # ---- row-sharded 1-vs-all: ComplEx, entity tables split by rows
    N, R, d, Q = 301, 5, 24, 9
    om, tabs = gpu.synthetic_case("complex", N, R, d, seed=7)
    rng = np.random.RandomState(3)
    qh, qr, qt = rng.randint(N, size=Q), rng.randint(R, size=Q), rng.randint(N, size=Q)
    ft, fh = gpu.random_filters_csr(rng, N, qh, qr, qt)
    want = oracle.rank_1vsall(om, qh, qr, qt, ft, fh)
    rlo, rhi = sharding.shard_range(N, world, rank)
    ent_local = [torch.from_numpy(tabs[0][rlo:rhi].copy()), torch.from_numpy(tabs[1][rlo:rhi].copy())]
    rel = [torch.from_numpy(tabs[2]), torch.from_numpy(tabs[3])]
    ranker = sharding.RowShardedRanker(_oracle_count_fn("complex", d, N), N, ent_local, rel, (0, 1), (2, 3))
    got = ranker.rank_queries(qh, qr, qt, ft, fh)
    np.testing.assert_array_equal(got.numpy(), want)
    # exchanged query rows are exact copies of the owners' rows
    uniq = np.unique(np.concatenate([qh, qt]))
    comp = ranker.exchange_query_rows(uniq)
    np.testing.assert_array_equal(comp[0].numpy(), tabs[0][uniq])
    np.testing.assert_array_equal(comp[1].numpy(), tabs[1][uniq])
    dist.barrier()
    dist.destroy_process_group()
    open(os.path.join(tmp, "ok%d" % rank), "w").write("ok")


def test_sharding_world2_gloo(tmp_path):
    world = 2
    port = _free_port()
    mp.start_processes(_worker, args=(world, port, str(tmp_path)), nprocs=world, join=True, start_method="spawn")
    assert all(os.path.exists(os.path.join(str(tmp_path), "ok%d" % r)) for r in range(world))


def test_shard_range_properties():
    from pykg2vec_b200.sharding import shard_range
    for total in (0, 1, 7, 8, 123182):
        for world in (1, 2, 3, 8):
            sizes = [shard_range(total, world, r)[1] - shard_range(total, world, r)[0] for r in range(world)]
            assert sum(sizes) == total and max(sizes) - min(sizes) <= 1
            assert shard_range(total, world, 0)[0] == 0 and shard_range(total, world, world - 1)[1] == total
