"""Helper for test_gpu_multi.py, launched with torch.distributed.run on 2+ GPUs: every rank builds one shard of one
logical index; the fused exchange (ab_comm: peer stores over NVLink + last-arriver merge) must equal the collective
route (NCCL all-gather + merge kernel) and the exact CPU merge of the shards' own lists, for raw and SQ8 spaces, and
the SQ shards must carry the single index's codes (global quantiser range)."""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from alayalite_b200 import _cabi  # noqa: E402
from alayalite_b200.sharded import ShardedIndex, global_quantizer_range, shard_range  # noqa: E402
from alayalite_b200.synthetic import low_rank  # noqa: E402

rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
torch.cuda.set_device(int(os.environ["LOCAL_RANK"]))
dev = torch.device("cuda", int(os.environ["LOCAL_RANK"]))
_cabi.check(_cabi.lib().ab_set_device(int(os.environ["LOCAL_RANK"])))
dist.init_process_group("nccl", device_id=dev)
n_total, dim, nq, k = 40000, 64, 501, 10
full, q = low_rank(n_total, dim, nq, 12, seed=9)
r0, r1 = shard_range(n_total, world, rank)
base = full[r0:r1]
ix = _cabi.NativeIndex("l2", "none", len(base), device=int(os.environ["LOCAL_RANK"]))
ix.fit(base, 80)
stream = torch.cuda.Stream(device=dev)
torch.cuda.set_stream(stream)
d_q = torch.from_numpy(q).to(dev)
fused = ShardedIndex(ix, r0, world, peer_memory=True)
coll = ShardedIndex(ix, r0, world, peer_memory=False)
ok = True
for it in range(4):  # several steps: exercises the two buffer parities
    for ef in (16, 64):
        a_ids, a_dist = [t.clone() for t in fused.batch_search_device(d_q, k, ef, stream)]
        b_ids, b_dist = [t.clone() for t in coll.batch_search_device(d_q, k, ef, stream.cuda_stream)]
        torch.cuda.synchronize()
        ok &= bool(torch.equal(a_ids, b_ids) and torch.equal(a_dist, b_dist))
        ok &= a_ids.shape[0] == len(fused.owned(nq))
# the full merged result on every rank, and what an exact merge of the shards' own lists gives
got = fused.batch_search(q, k, 64, device=dev, stream=stream)
ids = torch.empty((nq, k), dtype=torch.int32, device=dev)
dd = torch.empty((nq, k), dtype=torch.float32, device=dev)
ix.batch_search_device(d_q.data_ptr(), nq, k, 64, ids.data_ptr(), dd.data_ptr(), True, stream.cuda_stream)
ids += r0
gi = [torch.empty_like(ids) for _ in range(world)]
gd = [torch.empty_like(dd) for _ in range(world)]
dist.all_gather(gi, ids)
dist.all_gather(gd, dd)
torch.cuda.synchronize()
ci, cd = torch.cat(gi, 1).cpu().numpy().astype(np.int64), torch.cat(gd, 1).cpu().numpy()
order = np.lexsort((ci, cd), axis=1)[:, :k]
want = np.take_along_axis(ci, order, 1)
ok &= bool(np.array_equal(want, got.astype(np.int64)))
ok &= bool(np.array_equal(want[rank::world], a_ids.cpu().numpy().astype(np.int64)))
# quantised space: global range at fit time, final lists from the re-rank kernel
mn, mx = global_quantizer_range("l2", base)
ix8 = _cabi.NativeIndex("l2", "sq8", len(base), device=int(os.environ["LOCAL_RANK"]))
_cabi.fit_ranged(ix8, base, ef_construction=80, mn=mn, mx=mx)
wmn, wmx, wcodes = _cabi.quant_fit_encode("sq8", full)  # the single index's quantiser
smn, smx, scodes = ix8.codes()
ok &= bool(np.array_equal(smn, wmn) and np.array_equal(smx, wmx) and np.array_equal(scodes, wcodes[r0:r1]))
fused8 = ShardedIndex(ix8, r0, world, peer_memory=True)
coll8 = ShardedIndex(ix8, r0, world, peer_memory=False)
for ef in (32, 100):
    p_ids, p_dist = [t.clone() for t in fused8.batch_search_device(d_q, k, ef, stream)]
    c_ids, c_dist = [t.clone() for t in coll8.batch_search_device(d_q, k, ef, stream)]
    torch.cuda.synchronize()
    ok &= bool(torch.equal(p_ids, c_ids) and torch.equal(p_dist, c_dist))
errs = (fused._comm.errors() if fused._comm else -1, fused8._comm.errors() if fused8._comm else -1)
print(f"rank {rank} ok={ok} exchange={fused.exchange_kind!r} err={fused.peer_error} comm_errors={errs}", flush=True)
flag = torch.tensor([int(ok and fused.exchange_kind.startswith('fused') and errs == (0, 0))], device=dev)
dist.all_reduce(flag, op=dist.ReduceOp.MIN)
dist.destroy_process_group()
sys.exit(0 if int(flag.item()) == 1 else 1)
