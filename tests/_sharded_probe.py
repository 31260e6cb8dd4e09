"""Helper for test_gpu_multi.py, launched with torch.distributed.run on 2+ GPUs: every rank builds one shard,
then the merged results of the peer-memory exchange must equal those of the NCCL exchange and the exact merge."""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from alayalite_b200 import _cabi  # noqa: E402
from alayalite_b200.sharded import ShardedIndex  # noqa: E402
from alayalite_b200.synthetic import low_rank  # noqa: E402

rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
torch.cuda.set_device(int(os.environ["LOCAL_RANK"]))
dev = torch.device("cuda", int(os.environ["LOCAL_RANK"]))
dist.init_process_group("nccl", device_id=dev)
n, dim, nq, k = 20000, 64, 500, 10
base, q = low_rank(n, dim, nq, 12, seed=9, shard=rank)
ix = _cabi.NativeIndex("l2", "none", n)
ix.fit(base, 80)
stream = torch.cuda.Stream(device=dev)
torch.cuda.set_stream(stream)
d_q = torch.from_numpy(q).to(dev)
peer = ShardedIndex(ix, rank * n, world, peer_memory=True)
nccl = ShardedIndex(ix, rank * n, world, peer_memory=False)
ok = True
for it in range(4):  # several steps: exercises the two buffer parities
    for ef in (16, 64):
        a_ids, a_dist = [t.clone() for t in peer.batch_search_device(d_q, k, ef, stream.cuda_stream)]
        b_ids, b_dist = [t.clone() for t in nccl.batch_search_device(d_q, k, ef, stream.cuda_stream)]
        torch.cuda.synchronize()
        ok &= bool(torch.equal(a_ids, b_ids) and torch.equal(a_dist, b_dist))
# host route: pinned buffers are used in place, pageable ones are copied; same answer either way
qp = torch.from_numpy(q).pin_memory()
outp = torch.full((nq, k), -1, dtype=torch.int32).pin_memory()
peer.batch_search(qp, k, 64, out=outp, device=dev, stream=stream.cuda_stream)
got_np = peer.batch_search(q, k, 64, device=dev, stream=stream.cuda_stream)
ok &= bool(np.array_equal(outp.numpy(), a_ids.cpu().numpy()) and np.array_equal(got_np.view(np.int32), a_ids.cpu().numpy()))
# every rank holds the same merged list
g = [torch.empty_like(a_ids) for _ in range(world)]
dist.all_gather(g, a_ids)
ok &= all(bool(torch.equal(g[0], x)) for x in g)
# and it is what an exact merge of the shards' own lists gives
ids = torch.empty((nq, k), dtype=torch.int32, device=dev)
dd = torch.empty((nq, k), dtype=torch.float32, device=dev)
ix.batch_search_device(d_q.data_ptr(), nq, k, 64, ids.data_ptr(), dd.data_ptr(), True, stream.cuda_stream)
ids += rank * n
gi = [torch.empty_like(ids) for _ in range(world)]
gd = [torch.empty_like(dd) for _ in range(world)]
dist.all_gather(gi, ids)
dist.all_gather(gd, dd)
torch.cuda.synchronize()
ci, cd = torch.cat(gi, 1).cpu().numpy().astype(np.int64), torch.cat(gd, 1).cpu().numpy()
order = np.lexsort((ci, cd), axis=1)[:, :k]
want = np.take_along_axis(ci, order, 1)
ok &= bool(np.array_equal(want, a_ids.cpu().numpy().astype(np.int64)))
# quantised space: the final lists come from the re-rank kernel, which carries the peer stores there
ix8 = _cabi.NativeIndex("l2", "sq8", n)
ix8.fit(base, 80)
peer8 = ShardedIndex(ix8, rank * n, world, peer_memory=True)
nccl8 = ShardedIndex(ix8, rank * n, world, peer_memory=False)
for ef in (32, 100):
    p_ids, p_dist = [t.clone() for t in peer8.batch_search_device(d_q, k, ef, stream.cuda_stream)]
    c_ids, c_dist = [t.clone() for t in nccl8.batch_search_device(d_q, k, ef, stream.cuda_stream)]
    torch.cuda.synchronize()
    ok &= bool(torch.equal(p_ids, c_ids) and torch.equal(p_dist, c_dist))
print(f"rank {rank} ok={ok} exchange={peer.exchange_kind!r} err={getattr(peer, 'peer_error', None)}", flush=True)
flag = torch.tensor([int(ok and peer.exchange_kind.startswith('peer'))], device=dev)
dist.all_reduce(flag, op=dist.ReduceOp.MIN)
dist.destroy_process_group()
sys.exit(0 if int(flag.item()) == 1 else 1)
