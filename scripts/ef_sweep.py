"""Developer probe: device-resident traversal time per ef on the bench workload (1M x 128, L=16), CUDA events,
L2 flushed between launches.  python scripts/ef_sweep.py [ef ...]"""
import os, sys, time
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from alayalite_b200 import _cabi
from alayalite_b200.synthetic import low_rank
efs = [int(x) for x in sys.argv[1:]] or [20, 52, 100]
n, dim, nq = 1_000_000, 128, 10_000
base, q = low_rank(n, dim, nq, 16, seed=42)
ix = _cabi.NativeIndex("l2", "none", n)
t = time.time(); ix.fit(base, 100); print(f"fit {time.time()-t:.2f}s", flush=True)
gt, _ = _cabi.bruteforce_topk("l2", base, q, 10)
stream = torch.cuda.Stream()
d_q = torch.from_numpy(q).cuda(); d_ids = torch.zeros((nq, 10), dtype=torch.int32, device="cuda"); d_dist = torch.zeros((nq, 10), device="cuda")
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
with torch.cuda.stream(stream):
    for ef in efs:
        ts = []
        for it in range(8):
            flush.zero_()
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record(stream)
            ix.batch_search_device(d_q.data_ptr(), nq, 10, ef, d_ids.data_ptr(), d_dist.data_ptr(), True, stream.cuda_stream)
            b.record(stream); b.synchronize()
            if it >= 3: ts.append(a.elapsed_time(b))
        ids = d_ids.cpu().numpy().astype(np.uint32)
        rec = np.mean([len(set(ids[i]) & set(gt[i])) / 10 for i in range(nq)])
        st = ix.stats()
        bq = ((st["ndc"] + st["overlay"]) * 512 + st["hops"] * 128) / nq
        ms = float(np.mean(ts))
        print(f"ef={ef:4d} {ms:7.3f} ms  qps={nq/ms*1e3:10.0f} recall={rec:.4f} B/q={bq/1e3:6.1f}KB alg={nq/ms*1e3*bq/1e12:5.2f} TB/s frac={nq/ms*1e3*bq/6555.8e9:.3f} spilled={st['spilled']}", flush=True)
