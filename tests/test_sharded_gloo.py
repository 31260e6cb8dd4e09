"""The N>1 host logic (id offsets, all-gather layout, merge contract) on CPU: world_size 2, gloo.
The two device steps are injected: per-shard search = the oracle's exact top-k on the shard,
merge = a numpy (distance, id) merge with the contract of ab_merge_topk_device."""

import os
import socket

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, out_dir):
    import sys

    sys.path.insert(0, ROOT)
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from alayalite_b200.sharded import ShardedIndex
    from alayalite_b200.synthetic import low_rank
    from oracle.oracle import Checker

    orc = Checker("orc")
    n_shard, dim, nq, k = 700, 16, 25, 10
    base, _ = low_rank(n_shard, dim, 1, latent=5, seed=100 + rank)
    _, q = low_rank(n_shard, dim, nq, latent=5, seed=100)

    def search_fn(d_queries, nq_, topk, ef, d_ids, d_dist, stream):
        ids, dd = orc.bruteforce_topk("l2", d_queries.numpy(), base, topk)
        d_ids.copy_(torch.from_numpy(ids.astype(np.int32)))
        d_dist.copy_(torch.from_numpy(dd))

    def merge_fn(g_dist, g_ids, nq_, topk, out_ids, out_dist, stream):
        gd, gi = g_dist.numpy(), g_ids.numpy()  # [world, nq, k]
        for i in range(nq_):
            pairs = sorted((float(gd[w, i, j]), int(gi[w, i, j])) for w in range(gd.shape[0]) for j in range(topk))
            out_dist[i] = torch.tensor([p[0] for p in pairs[:topk]])
            out_ids[i] = torch.tensor([p[1] for p in pairs[:topk]], dtype=torch.int32)

    sh = ShardedIndex(None, rank * n_shard, world, merge_fn=merge_fn, search_fn=search_fn)
    got = sh.batch_search(q, k, 32, device="cpu")
    np.save(os.path.join(out_dir, f"ids{rank}.npy"), got)
    np.save(os.path.join(out_dir, f"base{rank}.npy"), base)
    if rank == 0:
        np.save(os.path.join(out_dir, "q.npy"), q)
    dist.barrier()
    dist.destroy_process_group()


def test_sharded_search_two_ranks_gloo(tmp_path, orc):
    world = 2
    mp.spawn(_worker, args=(world, _free_port(), str(tmp_path)), nprocs=world, join=True)
    ids = [np.load(tmp_path / f"ids{r}.npy") for r in range(world)]
    assert np.array_equal(ids[0], ids[1])  # every rank holds the merged result
    union = np.concatenate([np.load(tmp_path / f"base{r}.npy") for r in range(world)])
    q = np.load(tmp_path / "q.npy")
    want, _ = orc.bruteforce_topk("l2", q, union, 10)  # one unsharded exact search over the union
    assert np.array_equal(ids[0], want)
    assert ids[0].max() >= 700  # global ids of the second shard appear
