"""Developer probe: single-query / small-batch latency on 1M x 128 with the row-prefetch variants (AB_PREFETCH)."""
import os, sys
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench
from alayalite_b200.synthetic import low_rank_device
dev = torch.device("cuda", 0)
cfg = dict(bench.CONFIGS["c2"])
base, q = low_rank_device(cfg["n"], 128, 1000, 16, device=dev)
sdk, ix, _ = bench.make_index(cfg, base, 0)
del base
d_ids = torch.empty((1000, 10), dtype=torch.int32, device=dev); d_d = torch.empty((1000, 10), device=dev)
s = torch.cuda.Stream(); e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
for ef in (52, 100):
    for nq in (1, 32):
        for pf in ("1", "0", "1", "0", "2", "0", "1"):
            os.environ["AB_PREFETCH"] = pf
            ts = []
            for it in range(40):
                with torch.cuda.stream(s):
                    e0.record(s)
                    ix.batch_search_device(q[it:it + nq].data_ptr(), nq, 10, ef, d_ids.data_ptr(), d_d.data_ptr(), True, s.cuda_stream)
                    e1.record(s)
                e1.synchronize()
                ts.append(e0.elapsed_time(e1))
            print(f"ef={ef:3d} nq={nq:3d} AB_PREFETCH={pf}: {np.median(ts[5:])*1e3:6.1f} us", flush=True)
