"""Developer probe (not a bench line): search QPS / recall of the CUDA traversal on a reference-built
graph, next to the reference's own CPU search on the same graph. Usage:
    python scripts/quick_perf.py [n] [dim] [nq] [latent]
"""
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from alayalite_b200 import _cabi  # noqa: E402
from alayalite_b200.synthetic import low_rank  # noqa: E402
from oracle.oracle import Checker  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 200_000
dim = int(sys.argv[2]) if len(sys.argv) > 2 else 128
nq = int(sys.argv[3]) if len(sys.argv) > 3 else 10_000
latent = int(sys.argv[4]) if len(sys.argv) > 4 else 16
threads = os.cpu_count()
print(f"n={n} dim={dim} nq={nq} latent={latent} host threads={threads}", flush=True)
base, q = low_rank(n, dim, nq, latent)
ref = Checker("ref")
print("simd", ref.simd_level())
r = ref.index("l2", "none", n, dim)
t = time.time()
r.fit(base.copy(), 100, threads)
print(f"reference build {time.time() - t:.1f}s", flush=True)
g = r.graph()
ix = _cabi.NativeIndex("l2", "none", n)
t = time.time()
ix.fit_spaces(base)
ix.set_graph(g["adj"], g["levels"], g["lists_flat"], g["ep"])
print(f"upload {time.time() - t:.2f}s")
t = time.time()
gt, _ = _cabi.bruteforce_topk("l2", base, q, 10)
print(f"gpu brute force {time.time() - t:.2f}s", flush=True)


def rec(a):
    return np.mean([len(set(a[i]) & set(gt[i])) / 10 for i in range(len(a))])


for ef in (16, 32, 50, 64, 100, 200):
    ix.batch_search(q, 10, ef)
    t = time.time()
    reps = 3
    for _ in range(reps):
        ids = ix.batch_search(q, 10, ef)
    dt = (time.time() - t) / reps
    st = ix.stats()
    row = dim * 4
    bq = ((st["ndc"] + st["overlay"]) * row + st["hops"] * 128) / nq
    t = time.time()
    rid = r.batch_search_inplace(q, 10, ef, threads, solo=True)
    dtc = time.time() - t
    print(f"ef={ef:4d} gpu qps={nq / dt:10.0f} recall={rec(ids):.4f} ndc/q={st['ndc'] / nq:7.1f} hops/q={st['hops'] / nq:6.1f} "
          f"ov/q={st['overlay'] / nq:5.1f} B/q={bq / 1e3:7.1f}KB hbm={nq / dt * bq / 1e9:7.1f}GB/s | cpu({threads}thr solo) qps={nq / dtc:8.0f} "
          f"recall={rec(rid):.4f} same={np.mean(np.all(rid == ids, axis=1)):.3f}", flush=True)

# ---- GPU-built index, same data: build time and recall at equal ef --------------------------------
ix2 = _cabi.NativeIndex("l2", "none", n)
t = time.time()
ix2.fit(base, 100)
print(f"GPU fit (spaces + graph) {time.time() - t:.2f}s", flush=True)
g2 = ix2.graph()
deg = (g2["adj"] != 0xFFFFFFFF).sum(1)
degr = (g["adj"] != 0xFFFFFFFF).sum(1)
print(f"avg level-0 degree gpu={deg.mean():.2f} ref={degr.mean():.2f}; zero-degree gpu={(deg == 0).sum()} ; ep gpu={g2['ep']} ref={g['ep']} levels equal={np.array_equal(g2['levels'], g['levels'])}")
indeg = np.bincount(g2["adj"][g2["adj"] != 0xFFFFFFFF], minlength=n)
print(f"unreachable (in-degree 0) gpu={(indeg == 0).sum()} ref={(np.bincount(g['adj'][g['adj'] != 0xFFFFFFFF], minlength=n) == 0).sum()}")
for ef in (16, 32, 50, 64, 100, 200):
    a = ix2.batch_search(q, 10, ef)
    t = time.time()
    a = ix2.batch_search(q, 10, ef)
    dt = time.time() - t
    st = ix2.stats()
    b = ix.batch_search(q, 10, ef)
    print(f"ef={ef:4d} recall gpu-built={rec(a):.4f} ref-built={rec(b):.4f} gpu-built qps={nq / dt:9.0f} ndc/q={st['ndc'] / nq:7.1f}", flush=True)
