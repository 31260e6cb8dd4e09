"""Profiling target: build one index on the GPU and run a few device-resident searches (for ncu).
python scripts/prof_search.py --n 1000000 --dim 128 --latent 32 --quant none --metric l2 --topk 10 --ef 280 [--reps 3]"""
import argparse, os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from alayalite_b200 import _cabi
from alayalite_b200.synthetic import low_rank_device
ap = argparse.ArgumentParser()
ap.add_argument("--n", type=int, default=1_000_000); ap.add_argument("--dim", type=int, default=128)
ap.add_argument("--quant", default="none"); ap.add_argument("--metric", default="l2")
ap.add_argument("--latent", type=int, default=16); ap.add_argument("--topk", type=int, default=10)
ap.add_argument("--nq", type=int, default=10_000); ap.add_argument("--ef", type=int, default=64); ap.add_argument("--reps", type=int, default=3)
a = ap.parse_args()
dev = torch.device("cuda", 0)
base, q = low_rank_device(a.n, a.dim, a.nq, a.latent, seed=42, normalize=(a.metric != "l2"), device=dev)
ix = _cabi.NativeIndex(a.metric, a.quant, a.n)
ix.fit_ptr(base.data_ptr(), a.n, a.dim, 100)
del base
d_ids = torch.zeros((a.nq, a.topk), dtype=torch.int32, device=dev); d_dist = torch.zeros((a.nq, a.topk), device=dev)
s = torch.cuda.Stream()
torch.cuda.synchronize()
torch.cuda.profiler.start()  # ncu --profile-from-start off: only the searches below are seen (the build launches the same kernel)
for _ in range(a.reps):
    ix.batch_search_device(q.data_ptr(), a.nq, a.topk, a.ef, d_ids.data_ptr(), d_dist.data_ptr(), True, s.cuda_stream)
torch.cuda.synchronize()
print(ix.stats())
