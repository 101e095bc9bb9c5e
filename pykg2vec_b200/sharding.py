"""Multi-GPU partitioning of the hot path (one process per GPU, torch.distributed).

The reference is single-process/single-device; what is added here is the
minimum the path needs:

* 1-vs-all evaluation with replicated tables: test triples are independent
  (pykg2vec/utils/evaluator.py:313) -> contiguous query shards, NO collective in the data
  path, one final all-gather of the Q x 4 int32 ranks.
* 1-vs-all evaluation with ROW-SHARDED entity tables (tables that outgrow one GPU): every
  rank sweeps its rows for all queries; the only exchanges are (1) an all-gather of the query
  rows, each contributed by the rank that owns it (owners hold contiguous slices of the sorted
  unique query ids, so the gathered blocks concatenate into the compact query table), and
  (2) ONE all-gather of the partial rank counts [Q,4] int32, summed locally.
* data-parallel training with replicated tables (pykg2vec_b200/trainer.py): "grads" — every rank
  scores its own batch shard forward + backward into dense gradient buffers, one all-reduce per
  table, identical dense optimizer step everywhere; "ids" — for tiny batches (24 KB of ids at
  B=512) the ranks all-gather the ids instead and apply the global batch.

Everything here is host logic over torch.distributed and runs unchanged on the gloo
backend (tests/test_sharding_gloo.py, world_size 2, CPU) — the per-shard counting itself is
injected (`count_fn`) so those tests can plug the oracle in; the product always passes the
CUDA kernel (`cuda_count_fn`).
"""
import os

import numpy as np
import torch
import torch.distributed as dist


def init_distributed(backend=None):
    """Initialise from the torchrun environment (RANK / WORLD_SIZE / MASTER_*)."""
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if world == 1:
        return 0, 1
    if not dist.is_initialized():
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        # NCCL_DEBUG is left exactly as the launcher set it (the driver reads the communicator's rank
        # count from NCCL's INFO lines); bench.py keeps its one JSON line clean by pointing fd 1 at
        # stderr while the run is in progress.
        if backend is None:
            backend = "nccl" if torch.cuda.is_available() else "gloo"
        dist.init_process_group(backend=backend)
    return dist.get_rank(), dist.get_world_size()


def shard_range(total, world, rank):
    """Contiguous balanced partition of range(total): sizes differ by at most one."""
    base, rem = divmod(int(total), int(world))
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def allgather_batch_ids(ids):
    """ids: [k, B] int64 on the rank's device -> [k, world*B]: the global batch every rank
    trains on (replicated update)."""
    return allgather_batch_ids_async(ids)()


def allgather_batch_ids_async(ids):
    """Start the id all-gather (NCCL runs it on its own stream) and return a closure that makes
    the current stream wait for it and yields the [k, world*B] global batch.  Issue it BEFORE
    enqueueing independent work (the evaluation sweep) so the exchange hides behind it."""
    if not dist.is_initialized() or dist.get_world_size() == 1:
        return lambda: ids
    world = dist.get_world_size()
    k, b = ids.shape
    out = torch.empty((world * k, b), dtype=ids.dtype, device=ids.device)  # concatenated along dim 0
    work = dist.all_gather_into_tensor(out, ids.contiguous(), async_op=True)

    def finish():
        work.wait()
        return out.view(world, k, b).permute(1, 0, 2).reshape(k, world * b).contiguous()
    return finish


def gather_query_shards(local_counts, total):
    """local_counts: [q_local, 4] int32 of this rank's query shard (shard_range order) ->
    [total, 4] on every rank.  The single collective of replicated-table evaluation."""
    return gather_query_shards_async(local_counts, total)()


def gather_query_shards_async(local_counts, total):
    """Asynchronous form: returns a closure that waits and assembles the [total, 4] ranks."""
    if not dist.is_initialized() or dist.get_world_size() == 1:
        return lambda: local_counts
    world = dist.get_world_size()
    cap = -(-int(total) // world)
    buf = torch.zeros((cap, 4), dtype=local_counts.dtype, device=local_counts.device)
    buf[:local_counts.shape[0]] = local_counts
    out = torch.empty((world * cap, 4), dtype=local_counts.dtype, device=local_counts.device)
    work = dist.all_gather_into_tensor(out, buf, async_op=True)

    def finish():
        work.wait()
        o = out.view(world, cap, 4)
        parts = []
        for r in range(world):
            lo, hi = shard_range(total, world, r)
            parts.append(o[r, :hi - lo])
        return torch.cat(parts, dim=0)
    return finish


def cuda_count_fn(name, dim, **spec_kw):
    """The product's per-shard counting: kge_rank_1vsall on the local row shard."""
    from . import _lib

    def fn(shard_tables, query_tables, row_lo, row_hi, qh_c, qr, qt_c, tgt_h, tgt_t, filt_t, filt_h):
        desc = _lib.ModelDesc(name, shard_tables, dim, **spec_kw)
        qdesc = _lib.ModelDesc(name, query_tables, dim, **spec_kw)
        return _lib.rank_1vsall(desc, qh_c, qr, qt_c, filt_t, filt_h, row_lo=row_lo, row_hi=row_hi,
                                query_desc=qdesc, tgt_h=tgt_h, tgt_t=tgt_t)
    return fn


class RowShardedRanker:
    """1-vs-all ranks over entity tables partitioned by rows across the ranks.

    entity_tables_local: this rank's rows [row_lo, row_hi) of every ENTITY table (list, in
    the model's C-ABI order restricted to entity tables); relation tables are replicated.
    ent_slots / rel_slots: positions of entity / relation tables in the C-ABI table list
    (e.g. ComplEx: ent_slots=(0,1), rel_slots=(2,3))."""

    def __init__(self, count_fn, num_ent, entity_tables_local, relation_tables, ent_slots, rel_slots,
                 rank=None, world=None):
        self.count_fn = count_fn
        self.num_ent = int(num_ent)
        self.rank = dist.get_rank() if rank is None else rank
        self.world = dist.get_world_size() if world is None else world
        self.row_lo, self.row_hi = shard_range(num_ent, self.world, self.rank)
        self.ent_local = list(entity_tables_local)
        self.rel = list(relation_tables)
        self.ent_slots, self.rel_slots = tuple(ent_slots), tuple(rel_slots)
        self._bufs = {}
        for t in self.ent_local:
            assert t.shape[0] == self.row_hi - self.row_lo, "local shard has the wrong row count"

    def _assemble(self, ent_tables):
        n = len(self.ent_slots) + len(self.rel_slots)
        out = [None] * n
        for s, t in zip(self.ent_slots, ent_tables):
            out[s] = t
        for s, t in zip(self.rel_slots, self.rel):
            out[s] = t
        return out

    def _owner_layout(self, uniq):
        """Owners of the SORTED global ids `uniq`: rank g owns uniq[starts[g]:starts[g+1]] (its row range);
        cap = the largest block (the all-gather's common block size)."""
        bounds = [shard_range(self.num_ent, self.world, g) for g in range(self.world)]
        starts = np.array([np.searchsorted(uniq, lo) for lo, _ in bounds] + [len(uniq)], dtype=np.int64)
        cap = max(int((starts[1:] - starts[:-1]).max()), 1)
        return starts, cap

    def _gather_owner_blocks(self, mine, cap):
        """One all-gather per entity table of the owners' rows: -> per table a [world * cap, d] buffer whose
        block g holds rank g's rows (block-padded: rows beyond a rank's count are never addressed)."""
        out = []
        for t in self.ent_local:
            key = (cap, t.shape[1], t.dtype, t.device)
            bufs = self._bufs.get(key)
            if bufs is None:   # (re)used across calls: no allocation, no zero fill on the path
                bufs = (torch.empty((cap, t.shape[1]), dtype=t.dtype, device=t.device),
                        [torch.empty((self.world * cap, t.shape[1]), dtype=t.dtype, device=t.device)
                         for _ in self.ent_local])
                self._bufs[key] = bufs
            send, recvs = bufs
            if mine.numel():
                torch.index_select(t, 0, mine, out=send[:mine.numel()])
            recv = recvs[len(out)]
            if self.world == 1:
                recv.copy_(send)
            else:
                dist.all_gather_into_tensor(recv, send)
            out.append(recv)
        return out

    def exchange_query_rows(self, uniq):
        """Compact table of the rows `uniq` (SORTED global ids) of every entity table, identical on all
        ranks.  Rank g owns the contiguous slice of `uniq` that falls in its row range, so one
        all-gather per table of the owners' blocks (padded to the largest block) holds every row exactly
        once — no zero-padded sum.  (rank_queries addresses the gathered blocks in place; this form
        concatenates them into the compact [len(uniq), d] table.)"""
        dev = self.ent_local[0].device
        uniq = np.ascontiguousarray(uniq, dtype=np.int64)
        starts, cap = self._owner_layout(uniq)
        mine = torch.as_tensor(uniq[starts[self.rank]:starts[self.rank + 1]] - self.row_lo, dtype=torch.long, device=dev)
        blocks = self._gather_owner_blocks(mine, cap)
        counts = (starts[1:] - starts[:-1]).tolist()
        return [torch.cat([r.view(self.world, cap, -1)[g, :counts[g]] for g in range(self.world)], dim=0) for r in blocks]

    def rank_queries(self, qh, qr, qt, filt_t=None, filt_h=None):
        """qh/qr/qt: numpy int64 global ids (identical on all ranks).  filt_*: CSR numpy
        (ptr, idx) with global ids or None.  Returns [Q,4] int32 global ranks on every rank.

        Host path: ONE copy carries every integer array of the call to the device; the query rows are addressed
        inside the gathered owner blocks (index = owner * cap + position in the owner's block), so nothing is
        concatenated or zero-filled; send / receive buffers are reused across calls."""
        dev = self.ent_local[0].device
        qh = np.ascontiguousarray(qh, dtype=np.int64)
        qr = np.ascontiguousarray(qr, dtype=np.int64)
        qt = np.ascontiguousarray(qt, dtype=np.int64)
        uniq = np.unique(np.concatenate([qh, qt]))
        starts, cap = self._owner_layout(uniq)

        def block_index(ids):   # position of each id's row in the [world * cap, d] gathered buffer
            pos = np.searchsorted(uniq, ids)
            owner = np.searchsorted(starts, pos, side="right") - 1
            return owner * cap + (pos - starts[owner])

        mine = uniq[starts[self.rank]:starts[self.rank + 1]] - self.row_lo
        parts = [block_index(qh), block_index(qt), qr, qh, qt, mine]
        for f in (filt_t, filt_h):
            if f is not None:
                parts += [np.ascontiguousarray(f[0], dtype=np.int64), np.ascontiguousarray(f[1], dtype=np.int64)]
        packed = torch.from_numpy(np.concatenate(parts)).to(dev)
        views, o = [], 0
        for a in parts:
            views.append(packed[o:o + len(a)])
            o += len(a)
        qh_c, qt_c, qr_d, qh_d, qt_d, mine_d = views[:6]
        rest = views[6:]
        ft = fh = None
        if filt_t is not None:
            ft, rest = (rest[0], rest[1]), rest[2:]
        if filt_h is not None:
            fh = (rest[0], rest[1])
        blocks = self._gather_owner_blocks(mine_d, cap)
        counts = self.count_fn(self._assemble(self.ent_local), self._assemble(blocks), self.row_lo,
                               self.row_hi, qh_c, qr_d, qt_c, qh_d, qt_d, ft, fh)
        if self.world > 1:
            # the single exchange of the sweep: ONE all-gather of the partial rank counts, summed locally
            # (exact: integer counts of bit-identical scores over disjoint row shards)
            allc = torch.empty((self.world,) + tuple(counts.shape), dtype=counts.dtype, device=counts.device)
            dist.all_gather_into_tensor(allc.view(self.world * counts.shape[0], counts.shape[1]), counts.contiguous())
            counts = allc.sum(dim=0, dtype=counts.dtype)
        return counts
