"""Developer probe: the tcgen05 brute-force path vs the exact CUDA-core path and the oracle."""
import os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from alayalite_b200 import _cabi
from alayalite_b200.synthetic import low_rank, integer_grid
from oracle.oracle import Checker
orc = Checker("orc")

def run(tc, metric, base, q, k):
    os.environ["AB_BRUTEFORCE_TC"] = "1" if tc else "0"
    t = time.time()
    ids, dist = _cabi.bruteforce_topk(metric, base, q, k)
    return ids, dist, time.time() - t

for name, (base, q), k in [("grid 3000x32", integer_grid(3000, 32, 300, seed=7, hi=12), 17),
                           ("lowrank 20000x64", low_rank(20000, 64, 500, latent=10, seed=17), 10),
                           ("odd 5000x37", low_rank(5000, 37, 130, latent=6, seed=5), 100),
                           ("lowrank 300000x128", low_rank(300000, 128, 2000, latent=16, seed=3), 10)]:
    for metric in ("l2", "ip"):
        a_ids, a_d, ta = run(False, metric, base, q, k)
        b_ids, b_d, tb = run(True, metric, base, q, k)
        same_ids = np.mean(a_ids == b_ids)
        print(f"{name:22s} {metric} k={k:3d} exact-kernel {ta*1e3:8.1f} ms  tc {tb*1e3:8.1f} ms  ids equal {same_ids:.6f}  dist equal {np.mean(a_d == b_d):.6f}", flush=True)
        if base.shape[0] <= 20000:
            o_ids, o_d = orc.bruteforce_topk(metric, q, base, k, threads=8)
            print(f"{'':22s} vs oracle: exact-kernel ids {np.mean(a_ids == o_ids):.6f}  tc ids {np.mean(b_ids == o_ids):.6f}")
# timing on the bench shape, device-resident
import torch
base, q = low_rank(1_000_000, 128, 10_000, 16, seed=42)
d_base, d_q = torch.from_numpy(base).cuda(), torch.from_numpy(q).cuda()
ids = torch.empty((10000, 10), dtype=torch.int32, device="cuda"); dist = torch.empty((10000, 10), dtype=torch.float32, device="cuda")
res = {}
for tc in (0, 1):
    os.environ["AB_BRUTEFORCE_TC"] = str(tc)
    for rep in range(3):
        torch.cuda.synchronize(); t = time.time()
        _cabi.bruteforce_topk_device("l2", d_base.data_ptr(), 1_000_000, 128, d_q.data_ptr(), 10_000, 10, ids.data_ptr(), dist.data_ptr())
        torch.cuda.synchronize(); dt = time.time() - t
    res[tc] = (ids.cpu().numpy().copy(), dist.cpu().numpy().copy())
    print(f"1M x 128 x 10k queries top-10 tc={tc}: {dt*1e3:.1f} ms  ({2*1e6*1e4*128/dt/1e12:.1f} TFLOP/s equivalent)", flush=True)
print("1M ids equal", np.mean(res[0][0] == res[1][0]), "dist equal", np.mean(res[0][1] == res[1][1]))
