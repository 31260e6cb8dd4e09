"""Developer probe: end-to-end QPS of the host routes (numpy pageable via the SDK, pinned via the C ABI) on 10M x 128."""
import os, sys, time
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench
cfg = dict(bench.CONFIGS["main"]); n = int(sys.argv[1]) if len(sys.argv) > 1 else cfg["n"]; cfg["n"] = n
from alayalite_b200.synthetic import low_rank_device
dev = torch.device("cuda", 0)
base, q = low_rank_device(n, 128, 10_000, 16, device=dev)
sdk, ix, _ = bench.make_index(cfg, base, 0)
del base
qh = q.cpu().numpy(); ef = int(sys.argv[2]) if len(sys.argv) > 2 else 68
d_ids = torch.empty((10_000, 10), dtype=torch.int32, device=dev); d_d = torch.empty((10_000, 10), device=dev)
ix.batch_search_device(q.data_ptr(), 10_000, 10, ef, d_ids.data_ptr(), d_d.data_ptr(), True, None); torch.cuda.synchronize()
want = d_ids.cpu().numpy().view(np.uint32)
for name, fn in (("numpy", lambda: sdk.batch_search(qh, 10, ef)),):
    for _ in range(3): r = fn()
    t = time.perf_counter()
    for _ in range(20): r = fn()
    dt = (time.perf_counter() - t) / 20
    print(f"{name}: {dt*1e3:.3f} ms per 10k queries = {1e4/dt/1e6:.2f} M QPS; ids equal device-resident run: {np.array_equal(r, want)}")
one = qh[0]
for ef1 in (52, 100):
    for _ in range(5): sdk.search(one, 10, ef1)
    t = time.perf_counter()
    for _ in range(200): sdk.search(one, 10, ef1)
    print(f"Index.search single query ef={ef1}: {(time.perf_counter()-t)/200*1e3:.3f} ms")
for nqs in (64, 512):
    t = time.perf_counter()
    for _ in range(50): sdk.batch_search(qh[:nqs], 10, 52)
    print(f"batch of {nqs}: {(time.perf_counter()-t)/50*1e3:.3f} ms")
