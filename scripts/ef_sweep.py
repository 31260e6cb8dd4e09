"""Developer probe: device-resident traversal time per ef, CUDA events, L2 flushed between launches.
python scripts/ef_sweep.py [--n N --dim D --quant none|sq8|sq4 --metric l2|ip --latent L --topk K] [ef ...]"""
import argparse, os, sys, time
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from alayalite_b200 import _cabi
from alayalite_b200.synthetic import low_rank
ap = argparse.ArgumentParser()
ap.add_argument("--n", type=int, default=1_000_000); ap.add_argument("--dim", type=int, default=128)
ap.add_argument("--quant", default="none"); ap.add_argument("--metric", default="l2")
ap.add_argument("--latent", type=int, default=16); ap.add_argument("--topk", type=int, default=10)
ap.add_argument("--nq", type=int, default=10_000)
ap.add_argument("--device-gen", action="store_true", help="generate rows on the GPU (bench.py's generator)")
ap.add_argument("--prefetch", default="", help="comma-separated AB_PREFETCH values to run every ef with")
ap.add_argument("efs", nargs="*", type=int)
a = ap.parse_args()
efs = a.efs or [20, 52, 100]
n, dim, nq, k = a.n, a.dim, a.nq, a.topk
ix = _cabi.NativeIndex(a.metric, a.quant, n)
if a.device_gen:
    from alayalite_b200.synthetic import low_rank_device
    d_base, d_qq = low_rank_device(n, dim, nq, a.latent, seed=42, normalize=(a.metric != "l2"), device="cuda")
    t = time.time(); ix.fit_ptr(d_base.data_ptr(), n, dim, 100); print(f"fit {time.time()-t:.2f}s", flush=True)
    g_ids = torch.empty((nq, k), dtype=torch.int32, device="cuda"); g_d = torch.empty((nq, k), device="cuda")
    _cabi.bruteforce_topk_device("l2" if a.metric == "l2" else "ip", d_base.data_ptr(), n, dim, d_qq.data_ptr(), nq, k, g_ids.data_ptr(), g_d.data_ptr(), None)
    torch.cuda.synchronize(); gt = g_ids.cpu().numpy().astype(np.uint32); q = d_qq.cpu().numpy(); del d_base
else:
    base, q = low_rank(n, dim, nq, a.latent, seed=42, normalize=(a.metric != "l2"))
    t = time.time(); ix.fit(base, 100); print(f"fit {time.time()-t:.2f}s", flush=True)
    gt, _ = _cabi.bruteforce_topk("l2" if a.metric == "l2" else "ip", base, q, k)
row = {"none": dim * 4, "sq8": dim, "sq4": (dim + 1) // 2}[a.quant]
stream = torch.cuda.Stream()
d_q = torch.from_numpy(q).cuda(); d_ids = torch.zeros((nq, k), dtype=torch.int32, device="cuda"); d_dist = torch.zeros((nq, k), device="cuda")
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
with torch.cuda.stream(stream):
    for ef, pf in [(e, p) for e in efs for p in (a.prefetch.split(",") if a.prefetch else [None])]:
        if ef < k: continue
        if pf is not None:
            os.environ["AB_PREFETCH"] = pf; print(f"AB_PREFETCH={pf}", end=" ")
        ts = []
        for it in range(8):
            flush.zero_()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(stream)
            ix.batch_search_device(d_q.data_ptr(), nq, k, ef, d_ids.data_ptr(), d_dist.data_ptr(), True, stream.cuda_stream)
            e1.record(stream); e1.synchronize()
            if it >= 3: ts.append(e0.elapsed_time(e1))
        ids = d_ids.cpu().numpy().astype(np.uint32)
        rec = np.mean([len(set(ids[i]) & set(gt[i])) / k for i in range(nq)])
        st = ix.stats()
        bq = ((st["ndc"] + st["overlay"]) * row + st["hops"] * 128 + (k * dim * 4 if a.quant != "none" else 0)) / nq
        ms = float(np.median(ts))
        print(f"ef={ef:4d} {ms:7.3f} ms (min {min(ts):.3f} max {max(ts):.3f}) ovf_ids/q={st['overflow_ids']/nq:.0f} big={st['bitmap']}  qps={nq/ms*1e3:10.0f} recall@{k}={rec:.4f} B/q={bq/1e3:6.1f}KB alg={nq/ms*1e3*bq/1e12:5.2f} TB/s frac={nq/ms*1e3*bq/6555.8e9:.3f} bitmap={st['spilled']}", flush=True)
