import sys, time, os
sys.path.insert(0, os.getcwd())
import numpy as np, torch
import bench
from pykg2vec_b200.evaluator import build_filter_csr
kg = bench.make_graph()
tr = bench.build(kg, torch.device("cuda", 0))
ev = tr.evaluator
hr_t, tr_h = kg.read_cache_data("hr_t"), kg.read_cache_data("tr_h")
steps = bench.make_batches(kg, 30, 0)
host = []
for ids, q in steps:
    ft = build_filter_csr([(int(h), int(r)) for h, r, t in q], hr_t)
    fh = build_filter_csr([(int(t), int(r)) for h, r, t in q], tr_h)
    host.append((ids, q, ft, fh))
for i in range(5):
    tr.train_batch(host[i][0]); ev.rank_triples(host[i][1][:,0], host[i][1][:,1], host[i][1][:,2], host[i][2], host[i][3])
torch.cuda.synchronize()
tt = te = 0.0
for i in range(5, 30):
    ids, q, ft, fh = host[i]
    t0 = time.perf_counter(); tr.train_batch(ids); t1 = time.perf_counter()
    ev.rank_triples(q[:,0], q[:,1], q[:,2], ft, fh); t2 = time.perf_counter()
    tt += t1 - t0; te += t2 - t1
print("train_batch %.1f us  rank_triples %.1f us" % (tt/25*1e6, te/25*1e6))
# inside rank_triples
import cProfile, pstats
pr = cProfile.Profile(); pr.enable()
for i in range(5, 30):
    ids, q, ft, fh = host[i]
    tr.train_batch(ids); ev.rank_triples(q[:,0], q[:,1], q[:,2], ft, fh)
pr.disable()
pstats.Stats(pr).sort_stats("cumulative").print_stats(14)
