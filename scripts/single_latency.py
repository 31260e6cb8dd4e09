"""Developer probe: latency of single-query search (PyIndex::search route) and small batches."""
import os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from alayalite_b200 import _cabi
from alayalite_b200.synthetic import low_rank
base, q = low_rank(1_000_000, 128, 1000, 16, seed=42)
ix = _cabi.NativeIndex("l2", "none", len(base)); ix.fit(base, 100)
for ef in (52, 100):
    ix.search(q[0], 10, ef)
    t = time.perf_counter()
    for i in range(200): ix.search(q[i], 10, ef)
    print(f"ef={ef} single query: {(time.perf_counter()-t)/200*1e6:.1f} us")
    for nb in (8, 64, 512):
        ix.batch_search(q[:nb], 10, ef)
        t = time.perf_counter()
        for i in range(20): ix.batch_search(q[:nb], 10, ef)
        dt = (time.perf_counter()-t)/20
        print(f"ef={ef} batch {nb}: {dt*1e6:.1f} us  ({nb/dt:.0f} q/s)")
