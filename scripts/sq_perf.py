"""Developer probe: SQ8 / SQ4 search at GIST/RAG-like shapes: recall after the reference-style re-rank, QPS, and
the reference CPU path on the same data.  python scripts/sq_perf.py n dim quant metric topk latent"""
import os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from alayalite_b200 import _cabi
from alayalite_b200.synthetic import low_rank
from oracle.oracle import Checker
n, dim, quant, metric, topk, latent = int(sys.argv[1]), int(sys.argv[2]), sys.argv[3], sys.argv[4], int(sys.argv[5]), int(sys.argv[6])
nq = 10000
threads = os.cpu_count()
t = time.time(); base, q = low_rank(n, dim, nq, latent, seed=42, normalize=(metric != "l2")); print(f"gen {time.time()-t:.1f}s", flush=True)
ix = _cabi.NativeIndex(metric, quant, n)
t = time.time(); ix.fit(base, 100); print(f"GPU fit {time.time()-t:.2f}s", flush=True)
t = time.time(); gt, _ = _cabi.bruteforce_topk("l2" if metric == "l2" else "ip", base, q, topk); print(f"GPU brute force {time.time()-t:.2f}s", flush=True)
def rec(a): return np.mean([len(set(a[i]) & set(gt[i])) / topk for i in range(len(a))])
row = dim if quant == "sq8" else (dim + 1) // 2
for ef in (max(topk, 32), max(topk, 64), 100, 200, 400):
    if ef < topk: continue
    ix.batch_search(q, topk, ef)
    t = time.time(); ids = ix.batch_search(q, topk, ef); dt = time.time() - t
    st = ix.stats()
    bq = ((st["ndc"] + st["overlay"]) * row + st["hops"] * 128 + topk * dim * 4) / nq
    print(f"ef={ef:4d} qps={nq/dt:9.0f} recall@{topk}={rec(ids):.4f} ndc/q={st['ndc']/nq:7.1f} hops/q={st['hops']/nq:6.1f} B/q={bq/1e3:7.1f}KB alg GB/s={nq/dt*bq/1e9:7.1f} spilled={st['spilled']}", flush=True)
if len(sys.argv) > 7:
    ref = Checker("ref")
    r = ref.index(metric, quant, n, dim)
    t = time.time(); r.fit(base.copy(), 100, threads); print(f"reference CPU fit {time.time()-t:.1f}s", flush=True)
    for ef in (max(topk, 64), 100, 200):
        t = time.time(); rid = r.batch_search_inplace(q[:2000], topk, ef, threads); dt = time.time() - t
        print(f"  cpu ef={ef} qps={2000/dt:8.0f} recall={np.mean([len(set(rid[i]) & set(gt[i])) / topk for i in range(2000)]):.4f}", flush=True)
