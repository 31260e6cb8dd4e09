"""BASELINE.json's configs as GPU parity cases, checked through size-independent properties:
exact ground truth agreement, monotone recall, sorted distances, save -> load round trips, shard
merge == unsharded search. Sizes are the configs' own where a test stays within ~20 s on a B200
(C2 full size; C3 / C4 / C5 scaled down in rows only, dimensions and top-k as specified)."""

import numpy as np
import pytest

from conftest import recall

pytestmark = pytest.mark.gpu


def _gt(cabi, metric, base, q, k):
    return cabi.bruteforce_topk("l2" if metric == "l2" else "ip", base, q, k)[0]


def test_config2_sift1m_shape_full_size(cabi, tmp_path):
    """C2: 1M x 128 fp32 L2 graph index, 10k queries, top-10 -- the bench workload."""
    from alayalite_b200.synthetic import low_rank

    base, q = low_rank(1_000_000, 128, 10_000, 16, seed=42)
    ix = cabi.NativeIndex("l2", "none", 1_000_000)
    ix.fit(base, 100)
    gt = _gt(cabi, "l2", base, q, 10)
    # exact ground truth: spot-check 20 queries against float64 numpy
    d = ((base.astype(np.float64)[None, :, :] - q[:20].astype(np.float64)[:, None, :]) ** 2).sum(-1)
    assert recall(gt[:20], np.argsort(d, axis=1)[:, :10]) >= 0.995
    recs = []
    for ef in (16, 32, 64, 128):
        ids, dist = ix.batch_search_with_distance(q, 10, ef)
        assert np.all(np.diff(dist, axis=1) >= 0)  # pool order = ascending distance
        assert np.all(ids < 1_000_000)
        recs.append(recall(ids, gt))
    assert recs == sorted(recs) and recs[-1] >= 0.99 and recs[2] >= 0.95
    # distances returned are the exact fp32 distances of the returned ids
    ids, dist = ix.batch_search_with_distance(q[:64], 10, 64)
    want = ((base[ids.reshape(-1)].astype(np.float64) - np.repeat(q[:64], 10, axis=0).astype(np.float64)) ** 2).sum(1).reshape(64, 10)
    assert np.allclose(dist, want, rtol=1e-4)
    # save -> load round trip (reference file format): identical answers
    ix.save(str(tmp_path / "g.index"), str(tmp_path / "raw.data"))
    jx = cabi.NativeIndex("l2", "none", 1)
    jx.load(str(tmp_path / "g.index"), str(tmp_path / "raw.data"))
    assert np.array_equal(jx.batch_search(q[:2000], 10, 64), ix.batch_search(q[:2000], 10, 64))
    st = ix.stats()
    assert st["queries"] == 2000 and st["hops"] >= 2000 * 64


@pytest.mark.parametrize("quant,metric,n,dim,k", [("sq8", "l2", 200_000, 960, 10), ("sq4", "ip", 300_000, 768, 100), ("sq4", "cos", 100_000, 768, 100)])
def test_config3_config4_quantised_shapes(cabi, quant, metric, n, dim, k):
    """C3 (GIST-960 shape, SQ8, top-10) and C4 (RAG 768-d, IP / cosine, SQ4, top-100), fewer rows."""
    from alayalite_b200.synthetic import low_rank

    base, q = low_rank(n, dim, 1000, 32, seed=5, normalize=(metric != "l2"))
    ix = cabi.NativeIndex(metric, quant, n)
    ix.fit(base, 100)
    gt = _gt(cabi, metric, base, q, k)
    r_lo = recall(ix.batch_search(q, k, max(k, 64)), gt)
    r_hi = recall(ix.batch_search(q, k, 4 * max(k, 64)), gt)
    assert r_hi >= r_lo and r_hi >= 0.8
    # the re-rank orders the returned ids by exact raw-space distance (index.hpp:450-467)
    ids = ix.batch_search(q[:50], k, 2 * max(k, 64))
    for i in range(50):
        rows = base[ids[i]].astype(np.float64)
        qq = q[i].astype(np.float64)
        d = ((rows - qq) ** 2).sum(1) if metric == "l2" else -(rows @ qq)
        assert np.all(np.diff(d) >= -1e-4 * np.maximum(1.0, np.abs(d[:-1])))
    # un-reranked ids of the SQ space are what the quantised traversal ranks first
    raw_ids, dist = ix.batch_search_with_distance(q[:50], k, 2 * max(k, 64))
    assert dist.shape == (0, k) and raw_ids.shape == (50, k)


def test_config5_sharded_merge_equals_unsharded(cabi):
    """C5 (Deep-100M shape, 96-d, sharded by id, top-10 merge): two shards emulated on one GPU.
    Exhaustive search per shard + device merge must equal one exhaustive search over the union, and the
    merged graph search must reach the same recall as an unsharded index at equal ef."""
    import torch

    from alayalite_b200.synthetic import low_rank

    n, dim, nq, k = 400_000, 96, 2000, 10
    base, q = low_rank(n, dim, nq, 24, seed=9)
    half = n // 2
    gt = _gt(cabi, "l2", base, q, k)
    dev = torch.device("cuda", 0)
    stream = torch.cuda.Stream(device=dev)  # one explicit stream for torch ops and the library's kernels
    torch.cuda.set_stream(stream)
    sp = stream.cuda_stream
    g_ids = torch.empty((2, nq, k), dtype=torch.int32, device=dev)
    g_dist = torch.empty((2, nq, k), dtype=torch.float32, device=dev)
    out_ids = torch.empty((nq, k), dtype=torch.int32, device=dev)
    out_dist = torch.empty((nq, k), dtype=torch.float32, device=dev)
    # (1) exhaustive per shard -> merge == exhaustive over the union
    for s in range(2):
        ids, dist = cabi.bruteforce_topk("l2", base[s * half:(s + 1) * half], q, k)
        g_ids[s] = torch.from_numpy((ids + s * half).astype(np.int32)).to(dev)
        g_dist[s] = torch.from_numpy(dist).to(dev)
    cabi.merge_topk_device(g_dist.data_ptr(), g_ids.data_ptr(), 2, nq, k, out_ids.data_ptr(), out_dist.data_ptr(), sp)
    torch.cuda.synchronize()
    assert np.array_equal(out_ids.cpu().numpy().astype(np.uint32), gt)
    # (2) graph search per shard -> merge
    d_q = torch.from_numpy(q).to(dev)
    shards = []
    for s in range(2):
        ix = cabi.NativeIndex("l2", "none", half)
        ix.fit(base[s * half:(s + 1) * half], 100)
        shards.append(ix)
    whole = cabi.NativeIndex("l2", "none", n)
    whole.fit(base, 100)
    for ef in (32, 64):
        for s, ix in enumerate(shards):
            ix.batch_search_device(d_q.data_ptr(), nq, k, ef, g_ids[s].data_ptr(), g_dist[s].data_ptr(), True, sp)
            g_ids[s] += s * half
        cabi.merge_topk_device(g_dist.data_ptr(), g_ids.data_ptr(), 2, nq, k, out_ids.data_ptr(), out_dist.data_ptr(), sp)
        torch.cuda.synchronize()
        merged = out_ids.cpu().numpy().astype(np.uint32)
        assert np.all(np.diff(out_dist.cpu().numpy(), axis=1) >= 0)
        r_sharded, r_whole = recall(merged, gt), recall(whole.batch_search(q, k, ef), gt)
        assert r_sharded >= r_whole - 0.005, (ef, r_sharded, r_whole)  # 2 x ef candidates are examined
