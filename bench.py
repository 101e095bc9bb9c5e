#!/usr/bin/env python
"""bench.py — scored triples/sec (train + 1-vs-all eval), FB15k-237 shape, TransE d=200.

    python bench.py --gpus N --steps K --warmup W            # this repo's CUDA path
    python bench.py --impl reference --steps K --warmup W    # the reference's CPU path (port)

Workload (BASELINE.json configs[1]): TransE, N=14,541 entities, R=237 relations, d=200,
L2 norm, hinge margin 5.0 (pykg2vec/hyperparams/TransE.yaml:10), SGD lr 0.01, batch 512,
neg_rate 1, synthetic FB15k-237-shaped graph (no dataset is obtainable offline), tables
xavier-uniform random-init.

One STEP = what the reference repeats on this path for one batch of each kind:
  * one training batch (Trainer.train_model_epoch body, pykg2vec/utils/trainer.py:269-300):
    512 positive + 512 negative triples scored, hinge loss, backward, SGD update
    -> 1,024 scored triples;
  * one evaluation batch (Evaluator.test, pykg2vec/utils/evaluator.py:309-334) of Q=512 test
    triples, each ranked 1-vs-all against every entity in both directions, raw + filtered
    -> 2*512*14,541 = 14,889,984 scored triples.
`value` = scored triples / second of the whole job with inputs resident in HBM;
`e2e` = the same through the host API (Trainer.train_batch + Evaluator.rank_triples:
host id buffers in, pinned H2D, kernels, D2H of loss and ranks) — copies inside the timing.

Multi-GPU (torchrun, one rank per GPU): weak scaling — every rank brings its own training
batch and its own 512 test triples.  Training is data-parallel with replicated tables: the
ranks all-gather their batch ids over NCCL and each applies the identical global update
(pykg2vec_b200/sharding.py); evaluation shards the test triples with no data-path
collective and all-gathers the Q x 4 ranks at the end of the step.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

WORKLOAD = dict(model="transe", dataset="fb15k_237", N=14541, R=237, d=200, l1=False, margin=5.0,
                lr=0.01, B=512, neg=1, Q=512)
ALG_BYTES_PER_CANDIDATE = 800  # TransE 1-vs-all eval streams one d*4-byte row per score
L2_FLUSH_BYTES = 256 << 20


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        with open(p) as f:
            d = json.load(f)
        return float(d["hbm_gbs"]), "measured (MEASURED_PEAKS.json)", float(d.get("sm_max_mhz", 1965.0))
    return 6650.0, "fallback (B200_PROFILING.md)", 1965.0


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled DURING the timed region."""
    Q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.rows, self.proc, self.index = [], None, index

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.Q,
                                          "--format=csv,noheader,nounits", "-lms", "100"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([x.strip() for x in line.split(",")])

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm = [float(r[0]) for r in self.rows if len(r) >= 6 and r[0].replace(".", "").isdigit()]
        mx = [float(r[1]) for r in self.rows if len(r) >= 6 and r[1].replace(".", "").isdigit()]
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = sorted({names[i] for r in self.rows if len(r) >= 6 for i in range(4) if r[2 + i] == "Active"})
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": reasons, "samples": len(sm)}


def make_graph():
    from pykg2vec_b200.synthetic import SyntheticKnowledgeGraph
    return SyntheticKnowledgeGraph.shaped_like(WORKLOAD["dataset"], seed=0)


def make_batches(kg, nsteps, rank, seed=1):
    """Per step: (pairwise batch of 6 id arrays, Q test triples).  Negatives: head or tail
    corrupted with p=0.5, uniform (pykg2vec/data/generator.py:73-95, 'uniform' sampling)."""
    w = WORKLOAD
    rng = np.random.RandomState(seed + 7919 * rank)
    train, test = kg.arrays["train"], kg.arrays["test"]
    out = []
    for s in range(nsteps):
        sel = rng.randint(len(train), size=w["B"])
        ph, pr, pt = train[sel, 0].copy(), train[sel, 1].copy(), train[sel, 2].copy()
        corrupt_tail = rng.random_sample(w["B"]) > 0.5
        rnd = rng.randint(w["N"], size=w["B"])
        nh = np.where(corrupt_tail, ph, rnd)
        nt = np.where(corrupt_tail, rnd, pt)
        qsel = (np.arange(w["Q"]) + (s * w["Q"] + rank * 4099)) % len(test)
        out.append(([ph, pr, pt, nh, pr.copy(), nt], test[qsel]))
    return out


def build(kg, device):
    import torch
    import pykg2vec_b200
    from pykg2vec_b200.synthetic import SyntheticConfig
    from pykg2vec_b200.trainer import Trainer
    w = WORKLOAD
    cfg = SyntheticConfig(kg, device=device, optimizer="sgd", learning_rate=w["lr"], margin=w["margin"],
                          hidden_size=w["d"], l1_flag=w["l1"], batch_size=w["B"], neg_rate=w["neg"])
    torch.manual_seed(2)
    model = pykg2vec_b200.import_model(w["model"])(**cfg.__dict__)
    tr = Trainer(model, cfg)
    tr.build_model()
    return tr


def run_cuda(args):
    import torch
    import torch.distributed as dist
    from pykg2vec_b200 import _lib, sharding
    from pykg2vec_b200.evaluator import build_filter_csr
    rank, world = sharding.init_distributed()
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    w = WORKLOAD
    kg = make_graph()
    tr = build(kg, dev)
    ev = tr.evaluator
    hr_t, tr_h = kg.read_cache_data("hr_t"), kg.read_cache_data("tr_h")
    total = args.warmup + args.steps
    steps = make_batches(kg, total, rank)
    host = []   # host-side inputs per step (for the e2e leg)
    devin = []  # device-resident inputs per step (for the HBM-resident leg)
    for ids, q in steps:
        ft = build_filter_csr([(int(h), int(r)) for h, r, t in q], hr_t)
        fh = build_filter_csr([(int(t), int(r)) for h, r, t in q], tr_h)
        host.append((ids, q, ft, fh))
        tod = lambda a: torch.from_numpy(np.ascontiguousarray(a, dtype=np.int64)).to(dev)
        devin.append((torch.stack([tod(a) for a in ids]), tod(q[:, 0]), tod(q[:, 1]), tod(q[:, 2]),
                      (tod(ft[0]), tod(ft[1])), (tod(fh[0]), tod(fh[1]))))
    desc = tr.model.kge_desc()
    scratch = tr._grad_scratch
    loss_buf = torch.zeros(1, dtype=torch.float32, device=dev)
    counts = torch.zeros((w["Q"], 4), dtype=torch.int32, device=dev)
    ws = torch.empty(max(_lib.rank_workspace_bytes(desc, w["Q"]), 16), dtype=torch.uint8, device=dev)
    flush = torch.empty(L2_FLUSH_BYTES, dtype=torch.uint8, device=dev)
    scored_per_step = world * (w["B"] * (1 + w["neg"]) + 2 * w["Q"] * w["N"])

    pending = []  # asynchronous rank gathers of earlier steps (drained before the timing stops)

    def resident_step(i):
        if graph_step is not None:
            return graph_step(i)
        ids, qh, qr, qt, ft, fh = devin[i]
        # multi-GPU: start the 24 KB id all-gather first, sweep this rank's test triples while it
        # is in flight, then train on the gathered global batch (no-op closures at world 1)
        get_ids = sharding.allgather_batch_ids_async(ids)
        counts.zero_()
        _lib.rank_1vsall(desc, qh, qr, qt, ft, fh, counts=counts, workspace=ws)
        if world > 1:
            while pending:
                pending.pop()()
            pending.append(sharding.gather_query_shards_async(counts.clone(), world * w["Q"]))
        gids = get_ids()
        _lib.train_pairwise_hinge_sgd(desc, scratch, gids[0], gids[1], gids[2], gids[3], gids[4], gids[5],
                                      w["margin"], w["lr"], loss_buf)

    def e2e_step(i):
        ids, q, ft, fh = host[i]
        if world > 1:
            # the host API has no multi-GPU trainer yet: ids are exchanged on the device while the
            # evaluation batch runs
            dids, _ = tr._to_device(ids)
            get_ids = sharding.allgather_batch_ids_async(torch.stack(dids))
            ranks = ev.rank_triples(q[:, 0], q[:, 1], q[:, 2], ft, fh)
            g = get_ids()
            _lib.train_pairwise_hinge_sgd(desc, scratch, g[0], g[1], g[2], g[3], g[4], g[5], w["margin"],
                                          w["lr"], loss_buf)
            loss = float(loss_buf.item())
            return loss, ranks
        pending_loss = tr.train_batch(ids, sync=False)   # H2D + kernels + D2H enqueued as one graph
        ranks = ev.rank_triples(q[:, 0], q[:, 1], q[:, 2], ft, fh)   # host staging overlaps it; syncs
        return float(pending_loss), ranks                # pinned loss of this step, read on the host

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # Single GPU: the resident step (1 train batch + 1 eval batch = 11 short kernels) is captured
    # ONCE as a CUDA graph reading from fixed device buffers; each timed step is one D2D copy of
    # that step's resident inputs into those buffers plus one graph replay.
    graph_step = None
    if world == 1 and args.graph:
        cap_t = max(x[4][1].numel() for x in devin)
        cap_h = max(x[5][1].numel() for x in devin)
        s_ids = torch.zeros_like(devin[0][0])
        s_q = [torch.zeros(w["Q"], dtype=torch.int64, device=dev) for _ in range(3)]
        s_tp, s_hp = torch.zeros(w["Q"] + 1, dtype=torch.int64, device=dev), torch.zeros(w["Q"] + 1, dtype=torch.int64, device=dev)
        s_ti, s_hi = torch.zeros(cap_t, dtype=torch.int64, device=dev), torch.zeros(cap_h, dtype=torch.int64, device=dev)

        def load_inputs(i):
            ids, qh, qr, qt, ft, fh = devin[i]
            s_ids.copy_(ids); s_q[0].copy_(qh); s_q[1].copy_(qr); s_q[2].copy_(qt)
            s_tp.copy_(ft[0]); s_hp.copy_(fh[0])
            s_ti[:ft[1].numel()].copy_(ft[1]); s_hi[:fh[1].numel()].copy_(fh[1])

        def body():
            counts.zero_()
            _lib.rank_1vsall(desc, s_q[0], s_q[1], s_q[2], (s_tp, s_ti), (s_hp, s_hi), counts=counts, workspace=ws)
            _lib.train_pairwise_hinge_sgd(desc, scratch, s_ids[0], s_ids[1], s_ids[2], s_ids[3], s_ids[4], s_ids[5],
                                          w["margin"], 0.0 if body.warm else w["lr"], loss_buf)
        body.warm = True
        load_inputs(0)
        side = torch.cuda.Stream(device=dev)
        side.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(side):
            body()  # un-captured warm-up with lr = 0 (tables untouched)
        torch.cuda.current_stream().wait_stream(side)
        torch.cuda.synchronize()
        body.warm = False
        g = torch.cuda.CUDAGraph()
        k0 = _lib.launch_count()
        with torch.cuda.graph(g):
            body()
        kernels_per_replay = _lib.launch_count() - k0  # our kernels inside one replay of the graph

        def graph_step(i):
            load_inputs(i)
            g.replay()

    def timed(fn, first, n, use_events):
        """n steps starting at index `first`; L2 flushed (untimed) before every step; returns ms."""
        tot = 0.0
        for i in range(first, first + n):
            flush.zero_()
            barrier()
            if use_events:
                a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                a.record()
                fn(i)
                while pending:
                    pending.pop()()
                b.record()
                torch.cuda.synchronize()
                tot += a.elapsed_time(b)
                if os.environ.get("KGE_BENCH_DEBUG"):
                    sys.stderr.write("rank %d step %d: %.3f ms\n" % (rank, i, a.elapsed_time(b)))
            else:
                t0 = time.perf_counter()
                fn(i)
                torch.cuda.synchronize()
                tot += (time.perf_counter() - t0) * 1e3
        return tot

    def max_over_ranks(ms):
        if world == 1:
            return ms
        t = torch.tensor([ms], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    for i in range(args.warmup):
        resident_step(i)
        if not args.lite:
            e2e_step(i)
    barrier()
    # sustained run (~1.5 s of back-to-back steps) so that nvidia-smi samples clocks UNDER LOAD
    t_end = time.perf_counter() + (0.3 if args.lite else 1.5)
    sustained_steps = 0
    torch.cuda.synchronize()
    t_s0 = time.perf_counter()
    while time.perf_counter() < t_end:
        for i in range(args.warmup, total):
            resident_step(i)
        torch.cuda.synchronize()
        sustained_steps += args.steps
    ms_sustained = (time.perf_counter() - t_s0) * 1e3 / max(sustained_steps, 1)
    barrier()
    # no cyclic-GC pauses inside the timed legs: at N > 1 a pause on ONE rank stalls every rank's
    # id all-gather (seen as a single 3 ms step among 0.43 ms ones, profiles/r1_bench_n4_v3.err)
    import gc
    gc.collect()
    gc.disable()
    launches0 = _lib.launch_count()
    ms_res = max_over_ranks(timed(resident_step, args.warmup, args.steps, True))
    launches = _lib.launch_count() - launches0
    if graph_step is not None:
        launches = kernels_per_replay * args.steps  # replays re-execute the captured kernels
    # warm-L2 back-to-back variant (tables stay in the 126 MB L2 between steps, as in a real epoch)
    barrier()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for i in range(args.warmup, total):
        resident_step(i)
    while pending:
        pending.pop()()
    b.record()
    torch.cuda.synchronize()
    ms_warm = max_over_ranks(a.elapsed_time(b))
    ms_e2e = max_over_ranks(timed(e2e_step, args.warmup, args.steps, False)) if not args.lite else float("nan")
    gc.enable()
    # dominant kernel: the 1-vs-all sweep.  Timed alone (raw counts, one direction per launch pair)
    # with CUDA events on the launching stream.
    ids, qh, qr, qt, ft, fh = devin[args.warmup]
    reps = max(args.steps, 5)
    for _ in range(3):
        _lib.rank_1vsall(desc, qh, qr, qt, None, None, counts=counts, workspace=ws, flags=_lib.RANK_TAIL_ONLY)
    torch.cuda.synchronize()
    sweep_ms = 0.0
    for _ in range(reps):
        flush.zero_()
        torch.cuda.synchronize()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        _lib.rank_1vsall(desc, qh, qr, qt, None, None, counts=counts, workspace=ws, flags=_lib.RANK_TAIL_ONLY)
        b.record()
        torch.cuda.synchronize()
        sweep_ms += a.elapsed_time(b)
    sweep_ms /= reps
    clocks = sampler.stop() if rank == 0 else None
    if rank != 0:
        return None
    peak, peak_src, sm_max = peaks()
    alg_bytes = ALG_BYTES_PER_CANDIDATE * w["Q"] * w["N"]
    achieved = alg_bytes / (sweep_ms * 1e-3) / 1e9
    # secondary, honest bound of the batched sweep: fp32 pipe (2 instr / element-pair tail, 3 head)
    lane_ops = w["Q"] * w["N"] * w["d"] * 2
    fp32_peak = 148 * 128 * sm_max * 1e6
    cpu = cpu_baseline(sample_train=10, sample_queries=6) if not args.lite else None
    line = {
        "metric": "scored triples/sec (train + 1-vs-all eval)", "value": scored_per_step * args.steps / (ms_res * 1e-3),
        "unit": "triples/s", "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": ms_res / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f32", "data": "synthetic (FB15k-237-shaped random graph, random-init tables)",
        "config": {"workload": "TransE L2 d=200 on FB15k-237 shape (N=14541, R=237): per step one train batch "
                               "B=512 neg=1 hinge(margin 5)+SGD and one 1-vs-all eval batch of Q=512 test triples "
                               "(head+tail, raw+filtered)",
                   "scored_triples_per_step_per_gpu": w["B"] * (1 + w["neg"]) + 2 * w["Q"] * w["N"],
                   "l2": "flushed before every timed step (256 MiB memset, untimed); tables (11.6 MB) "
                         "otherwise stay L2-resident",
                   "parallelism": "dp%d: batch ids all-gathered, replicated update; test triples sharded" % world},
        "ms_per_step_warm_l2": ms_warm / args.steps,
        "ms_per_step_sustained": ms_sustained,
        "e2e": {"value": scored_per_step * args.steps / (ms_e2e * 1e-3), "unit": "triples/s",
                "ms_per_step": ms_e2e / args.steps,
                "h2d_bytes_per_step": int(6 * w["B"] * 8 + 3 * w["Q"] * 8 +
                                          sum(x.nbytes for x in host[args.warmup][2]) + sum(x.nbytes for x in host[args.warmup][3])),
                "d2h_bytes_per_step": 4 + w["Q"] * 4 * 4},
        "gpu_launches": int(launches),
        "clocks": clocks,
        "roofline": {"kernel": "1-vs-all sweep (tail direction, Q=512 x N=14541)", "bound": "hbm",
                     "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                     "peak_source": peak_src,
                     # dram__bytes_read+write of one sweep_tiled_kernel launch, ncu --set full
                     # (profiles/r1_ncu_sweep_tiled_v2_summary.txt)
                     "traffic": 12088576, "launch_ms": sweep_ms,
                     "algorithmic_bytes_per_launch": alg_bytes,
                     "note": "algorithmic bytes = the reference-equivalent streaming formulation (one 800-byte "
                             "row per scored candidate, SURVEY.md 8d); a sweep that batches queries re-uses "
                             "rows on chip, so frac can exceed 1 and the binding unit is the fp32 pipe",
                     "fp32_pipe": {"lane_ops_per_launch": lane_ops, "achieved_tlops": lane_ops / (sweep_ms * 1e-3) / 1e12,
                                   "peak_tlops": fp32_peak / 1e12, "frac": lane_ops / (sweep_ms * 1e-3) / fp32_peak}},
        "cpu_baseline": cpu,
    }
    return line


# ------------------------------------------------------------------ CPU reference arm ----
def usable_cores():
    """cores this process may actually use: affinity mask capped by the cgroup CPU quota."""
    n = len(os.sched_getaffinity(0)) if hasattr(os, "sched_getaffinity") else (os.cpu_count() or 1)
    try:
        with open("/sys/fs/cgroup/cpu.max") as f:
            quota, period = f.read().split()
        if quota != "max":
            n = max(1, min(n, int(int(quota) / int(period))))
    except Exception:
        pass
    return n


class CpuArm:
    """The torch port of the reference's CPU path (oracle/ref_port.py: the same ATen op chain,
    dense autograd + dense optim.SGD, forward over N + topk(N) + Python rank walk) on the host
    cores.  Bench/test infrastructure only."""

    def __init__(self):
        import torch
        from oracle import ref_port
        w = WORKLOAD
        self.torch, self.rp = torch, ref_port
        self.cores = usable_cores()
        torch.set_num_threads(self.cores)
        self.kg = make_graph()
        gen = torch.Generator().manual_seed(2)
        self.ent = ref_port.xavier_uniform(w["N"], w["d"], gen).requires_grad_()
        self.rel = ref_port.xavier_uniform(w["R"], w["d"], gen).requires_grad_()
        self.opt = torch.optim.SGD([self.ent, self.rel], lr=w["lr"])
        self.batches = make_batches(self.kg, 64, 0)
        self.hr_t, self.tr_h = self.kg.read_cache_data("hr_t"), self.kg.read_cache_data("tr_h")
        self.cursor = 0

    def train_steps(self, n):
        torch, rp, w = self.torch, self.rp, WORKLOAD
        t0 = time.perf_counter()
        for _ in range(n):
            ids, _q = self.batches[self.cursor % len(self.batches)]
            self.cursor += 1
            tid = [torch.LongTensor(np.asarray(a)) for a in ids]  # trainer.py:288-293
            self.opt.zero_grad()
            pos = rp.score("transe", [self.ent, self.rel], tid[0], tid[1], tid[2], l1_flag=w["l1"])
            neg = rp.score("transe", [self.ent, self.rel], tid[3], tid[4], tid[5], l1_flag=w["l1"])
            loss = rp.pairwise_hinge(pos, neg, w["margin"])
            loss.backward()
            self.opt.step()
            loss.item()
        return (time.perf_counter() - t0) / n

    def eval_queries(self, n):
        torch, rp, w = self.torch, self.rp, WORKLOAD
        test = self.kg.arrays["test"]
        q = [tuple(int(x) for x in test[(self.cursor * 7 + k) % len(test)]) for k in range(n)]
        fn = lambda a, b, c: rp.score("transe", [self.ent, self.rel], a, b, c, l1_flag=w["l1"])
        with torch.no_grad():
            t0 = time.perf_counter()
            rp.evaluate(fn, w["N"], q, self.hr_t, self.tr_h)
            return (time.perf_counter() - t0) / n


def tune_threads(arm):
    """Give the CPU arm its best shot: intra-op thread counts up to the usable cores are
    probed on one train step + one test triple and the fastest is kept (torch CPU kernels
    on small tensors often run faster on fewer threads than cores)."""
    w = WORKLOAD
    cands = sorted({c for c in (arm.cores, 64, 32, 16, 8, 4) if c <= arm.cores}, reverse=True)
    best, best_t = cands[0], float("inf")
    for c in cands:
        arm.torch.set_num_threads(c)
        arm.train_steps(1)
        t0 = time.perf_counter()
        t = arm.train_steps(1) + w["Q"] * arm.eval_queries(1)
        if t < best_t:
            best, best_t = c, t
        if time.perf_counter() - t0 > 20:
            break
    arm.torch.set_num_threads(best)
    arm.cores = best
    return best


def cpu_measure(n_train, n_queries, arm=None):
    """(seconds per train step, seconds per test triple, cores) after a short warm-up."""
    arm = arm or CpuArm()
    tune_threads(arm)
    arm.train_steps(2)
    arm.eval_queries(1)
    return arm.train_steps(n_train), arm.eval_queries(n_queries), arm.cores


def cpu_line_value(t_train, t_query):
    w = WORKLOAD
    step_s = t_train + w["Q"] * t_query
    return (w["B"] * (1 + w["neg"]) + 2 * w["Q"] * w["N"]) / step_s, step_s


def cpu_baseline(sample_train, sample_queries):
    t_train, t_query, cores = cpu_measure(sample_train, sample_queries)
    value, step_s = cpu_line_value(t_train, t_query)
    return {"value": value, "unit": "triples/s", "cores": cores, "kind": "port",
            "sample": "%d train steps + %d test triples of the same workload, extrapolated to one step "
                      "(1 train batch + 512 test triples); torch %s CPU port of the reference op chain incl. "
                      "topk + Python rank walk" % (sample_train, sample_queries, __import__("torch").__version__),
            "train_step_ms": t_train * 1e3, "eval_ms_per_test_triple": t_query * 1e3}


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return None
    w = WORKLOAD
    # each step: a bounded sample (1 train step + 2 test triples), extrapolated to the step composition
    per_step_queries = 2
    tt, tq = [], []
    arm = CpuArm()
    cores = tune_threads(arm)
    for s in range(args.warmup + args.steps):
        a, b = arm.train_steps(1), arm.eval_queries(per_step_queries)
        if s >= args.warmup:
            tt.append(a)
            tq.append(b)
    t_train, t_query = float(np.mean(tt)), float(np.mean(tq))
    value, step_s = cpu_line_value(t_train, t_query)
    import torch
    line = {
        "impl": "reference", "metric": "scored triples/sec (train + 1-vs-all eval)", "value": value,
        "unit": "triples/s", "n_gpus": int(os.environ.get("WORLD_SIZE", "1")), "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": step_s * 1e3, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic (FB15k-237-shaped random graph, random-init tables)",
        "config": {"workload": "TransE L2 d=200 on FB15k-237 shape (N=14541, R=237): per step one train batch "
                               "B=512 neg=1 hinge(margin 5)+SGD and one 1-vs-all eval batch of Q=512 test triples "
                               "(head+tail, raw+filtered)"},
        "cpu_baseline": {"value": value, "unit": "triples/s", "cores": cores, "kind": "port",
                         "sample": "per step 1 train batch + %d test triples timed, extrapolated to 512 test "
                                   "triples; torch %s CPU port of the reference op chain (dense autograd + "
                                   "optim.SGD; forward over N + topk(N) + Python rank walk)" % (per_step_queries, torch.__version__),
                         "train_step_ms": t_train * 1e3, "eval_ms_per_test_triple": t_query * 1e3},
        "e2e": {"value": value, "unit": "triples/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    return line


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="cuda", choices=["cuda", "reference"])
    ap.add_argument("--graph", action="store_true",
                    help="replay the resident step as one CUDA graph (measured: no gain — the step is bound "
                         "by kernel time, not launch latency — so the default launches kernel by kernel)")
    ap.add_argument("--lite", action="store_true",
                    help="profiling aid: only the HBM-resident leg (no e2e / CPU baseline); never a bench value")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "cuda" else args.warmup
    # stdout carries exactly ONE JSON line: while the run is in progress fd 1 points at stderr so
    # that banners printed by native libraries (e.g. NCCL's version line) cannot end up there
    sys.stdout.flush()
    saved = os.dup(1)
    os.dup2(2, 1)
    line = None
    try:
        line = run_reference(args) if args.impl == "reference" else run_cuda(args)
    finally:
        sys.stdout.flush()
        os.dup2(saved, 1)
        os.close(saved)
    if line is not None:
        print(json.dumps(line), flush=True)
    try:
        import torch.distributed as dist
        if dist.is_available() and dist.is_initialized():
            dist.destroy_process_group()
    except Exception:
        pass


if __name__ == "__main__":
    main()
