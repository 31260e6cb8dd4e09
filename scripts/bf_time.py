"""Developer probe: wall / device time of the exact top-k (tcgen05 path) on 1M x 128 x 10k."""
import os, sys, time
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from alayalite_b200 import _cabi
from alayalite_b200.synthetic import low_rank_device
dev = torch.device("cuda", 0)
n, dim, nq, k = 1_000_000, 128, 10_000, 10
base, q = low_rank_device(n, dim, nq, 16, device=dev)
ids = torch.empty((nq, k), dtype=torch.int32, device=dev); dd = torch.empty((nq, k), device=dev)
for _ in range(3): _cabi.bruteforce_topk_device("l2", base.data_ptr(), n, dim, q.data_ptr(), nq, k, ids.data_ptr(), dd.data_ptr(), None)
torch.cuda.synchronize()
w = []
for _ in range(10):
    t = time.perf_counter(); _cabi.bruteforce_topk_device("l2", base.data_ptr(), n, dim, q.data_ptr(), nq, k, ids.data_ptr(), dd.data_ptr(), None); torch.cuda.synchronize(); w.append((time.perf_counter() - t) * 1e3)
ref = torch.cdist(q[:64].double(), base.double()).pow(2).topk(k, largest=False).indices
print(f"bruteforce wall ms median {np.median(w):.2f} min {min(w):.2f}; TF/s {2*nq*n*dim/np.median(w)/1e9:.0f}; ids==float64 on 64 queries: {bool(torch.equal(ref.int(), ids[:64]))}")
