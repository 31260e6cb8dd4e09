"""Developer probe: recall of GPU-built vs reference-built graphs at equal ef, for several batch schedules."""
import os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from alayalite_b200 import _cabi
from alayalite_b200.synthetic import low_rank
from oracle.oracle import Checker
ref = Checker("ref")
def rec(a, gt): return np.mean([len(set(a[i]) & set(gt[i])) / 10 for i in range(len(a))])
for (n, dim, lat) in [(200000, 128, 16), (1000000, 128, 16)]:
    base, q = low_rank(n, dim, 2000, lat, seed=17)
    gt, _ = _cabi.bruteforce_topk("l2", base, q, 10)
    r = ref.index("l2", "none", n, dim); t = time.time(); r.fit(base.copy(), 100, 16); tr = time.time() - t
    efs = (16, 40, 100)
    rr = [rec(r.batch_search(q, 10, ef, 16, solo=True), gt) for ef in efs]
    print(f"n={n} ref build {tr:.1f}s recall {np.round(rr, 4)}", flush=True)
    for div, maxb in [(32, 2048), (64, 2048), (64, 1024), (64, 512), (128, 256), (64, 4096)]:
        os.environ["AB_BUILD_DIV"] = str(div); os.environ["AB_BUILD_MAXB"] = str(maxb)
        ix = _cabi.NativeIndex("l2", "none", n); t = time.time(); ix.fit(base, 100); tb = time.time() - t
        gg = [rec(ix.batch_search(q, 10, ef), gt) for ef in efs]
        print(f"   div={div:3d} maxb={maxb:6d} build {tb:.2f}s recall {np.round(gg, 4)} delta {np.round(np.array(gg) - np.array(rr), 4)}", flush=True)
