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


# ---- data-parallel Trainer logic on gloo / CPU with a device double ---------------------------------
def _install_cpu_double():
    """Swap the C-ABI bindings the fused TransE steps use for torch-CPU restatements (oracle/ref_port.py), so
    that the HOST logic of pykg2vec_b200.trainer.Trainer — mode selection, id all-gather (sync and async),
    gradient all-reduce with sum / average semantics, dense optimizer bookkeeping — runs without a GPU."""
    from oracle import ref_port
    from pykg2vec_b200 import _lib
    _lib._dev_f32 = lambda t, name: t
    _lib._dev_i64 = lambda t, name: t

    def score(desc, h, r, t, tables=None):
        return ref_port.score(desc.name, tables if tables is not None else desc.tables, h, r, t, l1_flag=desc.l1_flag)

    def score_fwd(desc, h, r, t, grouping=0, out=None):
        return score(desc, h, r, t).detach()

    def score_bwd(desc, h, r, t, grad_scores, grad_tables):
        with torch.enable_grad():   # the Trainer runs its fused steps under no_grad
            tabs = [w.detach().clone().requires_grad_() for w in desc.tables]
            score(desc, h, r, t, tabs).backward(grad_scores)
        for g, w in zip(grad_tables, tabs):
            if g is not None and w.grad is not None:
                g += w.grad

    def loss_pairwise_hinge(pos, neg, margin, want_grad=True):
        v = torch.relu(pos + margin - neg)
        gp = (v > 0).float()
        return v.sum().reshape(1), gp, -gp

    def optim_apply_dense(w, g, optimizer, lr, state1=None, state2=None, eps=None, betas=(0.9, 0.999), step=1):
        if optimizer == 0:
            w -= lr * g
        elif optimizer == 1:
            state1 += g * g
            w -= lr * g / (state1.sqrt() + (1e-10 if eps is None else eps))
        else:
            state1 += (g - state1) * (1 - betas[0])
            state2.mul_(betas[1]).addcmul_(g, g, value=1 - betas[1])
            denom = state2.sqrt() / (1 - betas[1] ** step) ** 0.5 + (1e-8 if eps is None else eps)
            w -= (lr / (1 - betas[0] ** step)) * state1 / denom
        g.zero_()

    def optim_apply_rows(desc, grad_scratch, state, optimizer, h, r, t, lr, eps=1e-10):
        for k, w in enumerate(desc.tables):   # dense equivalent; a second call sees zeroed gradients
            if grad_scratch[k] is not None:
                optim_apply_dense(w, grad_scratch[k], optimizer, lr, state[k] if state else None)

    def train_pairwise_hinge_sgd(desc, grad_scratch, ph, pr, pt, nh, nr, nt, margin, lr, loss_out=None):
        pos, neg = score_fwd(desc, ph, pr, pt), score_fwd(desc, nh, nr, nt)
        loss, gp, gn = loss_pairwise_hinge(pos, neg, margin)
        score_bwd(desc, ph, pr, pt, gp, grad_scratch)
        score_bwd(desc, nh, nr, nt, gn, grad_scratch)
        optim_apply_rows(desc, grad_scratch, None, 0, ph, pr, pt, lr)
        if loss_out is not None:
            loss_out.copy_(loss)
            return loss_out
        return loss

    for name, fn in list(locals().items()):
        if callable(fn) and hasattr(_lib, name) and name not in ("score",):
            setattr(_lib, name, fn)


def _dp_worker(rank, world, port, tmp):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    import sys
    sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
    import pykg2vec_b200
    from pykg2vec_b200 import sharding
    from pykg2vec_b200.synthetic import SyntheticConfig, SyntheticKnowledgeGraph
    from pykg2vec_b200.trainer import Trainer
    sharding.init_distributed(backend="gloo")
    _install_cpu_double()
    kg = SyntheticKnowledgeGraph(120, 5, 600, 20, 20, seed=1)

    def make(opt, mode):
        cfg = SyntheticConfig(kg, device="cpu", optimizer=opt, learning_rate=0.05, hidden_size=16, margin=1.0,
                              l1_flag=False, batch_size=32, neg_rate=1, dp_mode=mode, cuda_graph=False)
        torch.manual_seed(0)
        tr = Trainer(pykg2vec_b200.import_model("transe")(**cfg.__dict__), cfg)
        tr.build_model()
        return tr

    B = 32
    # mode selection: a tiny batch against these tables exchanges ids; a batch that out-weighs the tables, gradients
    auto = make("sgd", None)
    assert auto._world == world and auto._dp in ("ids", "grads")
    auto.config.batch_size = 8          # 16 triples x 32 floats < 2,000 table floats
    assert auto._pick_dp_mode() == "ids"
    auto.config.batch_size = 4096
    assert auto._pick_dp_mode() == "grads"
    # optimizers that divide by a gradient magnitude never exchange ids (replicas would drift, trainer.py)
    for opt in ("adagrad", "adam"):
        t_ = make(opt, None)
        t_.config.batch_size = 8
        assert t_._dp == "grads" and t_._pick_dp_mode() == "grads", opt
    for opt in ("sgd", "adagrad", "adam"):
        single = make(opt, "off")
        trs = {mode: make(opt, mode) for mode in ("grads", "ids")}
        assert single._dp is None and all(t_._dp == m for m, t_ in trs.items())
        for step in range(3):
            per_rank = []
            for rk in range(world):
                rng = np.random.RandomState(100 * step + rk)
                per_rank.append([rng.randint(120, size=B), rng.randint(5, size=B), rng.randint(120, size=B),
                                 rng.randint(120, size=B), rng.randint(5, size=B), rng.randint(120, size=B)])
            to = lambda a: torch.from_numpy(np.ascontiguousarray(a, dtype=np.int64))
            glob = [to(np.concatenate([per_rank[rk][k] for rk in range(world)])) for k in range(6)]
            l_single = float(single.train_batch_device(glob).item())
            for mode, t_ in trs.items():
                mine = [to(a) for a in per_rank[rank]]
                ex = t_.exchange_batch_async(mine)          # None in "grads" mode
                assert (ex is None) == (mode == "grads")
                l_dp = float(t_.train_batch_device(mine, exchanged=ex).item())
                assert abs(l_dp - l_single) <= 1e-4 * max(abs(l_single), 1e-6), (opt, mode, step, l_dp, l_single)
        for mode, t_ in trs.items():
            for (ka, va), (kb, vb) in zip(t_.model.state_dict().items(), single.model.state_dict().items()):
                np.testing.assert_allclose(va.numpy(), vb.numpy(), rtol=0, atol=2e-5, err_msg="%s %s %s" % (opt, mode, ka))
                other = va.clone()
                dist.broadcast(other, src=0)
                assert torch.equal(other, va), (opt, mode, ka)     # replicas stay identical
    dist.barrier()
    dist.destroy_process_group()
    open(os.path.join(tmp, "dp_ok%d" % rank), "w").write("ok")


def test_data_parallel_trainer_world2_gloo(tmp_path):
    world = 2
    mp.start_processes(_dp_worker, args=(world, _free_port(), str(tmp_path)), nprocs=world, join=True, start_method="spawn")
    assert all(os.path.exists(os.path.join(str(tmp_path), "dp_ok%d" % r)) for r in range(world))
