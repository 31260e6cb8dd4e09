"""GPU tests against artefacts of the compiled reference: the committed golden fixtures
(tests/golden/ref_small.npz), the incremental-update semantics, and the on-disk formats
(our files load in the reference and vice versa; byte-identical where the inputs are)."""

import os

import numpy as np
import pytest

from conftest import recall

pytestmark = pytest.mark.gpu

HERE = os.path.dirname(os.path.abspath(__file__))
GOLD = np.load(os.path.join(HERE, "golden", "ref_small.npz"))
CONFIGS = [("l2", "none"), ("ip", "none"), ("cos", "none"), ("l2", "sq8"), ("ip", "sq8"), ("l2", "sq4"), ("cos", "sq4")]


def _on_golden_graph(cabi, metric, quant):
    tag = f"{metric}_{quant}"
    ix = cabi.NativeIndex(metric, quant, 640)
    ix.fit_spaces(GOLD["base"])
    ix.set_graph(GOLD[f"{tag}/adj"], GOLD[f"{tag}/levels"], GOLD[f"{tag}/lists"], int(GOLD[f"{tag}/ep"]))
    return tag, ix


@pytest.mark.parametrize("metric,quant", CONFIGS)
def test_gpu_reproduces_reference_golden(cabi, metric, quant):
    tag, ix = _on_golden_graph(cabi, metric, quant)
    q = GOLD["queries"]
    if quant != "none" and metric != "cos":
        mn, mx, codes = ix.codes()
        assert np.array_equal(mn, GOLD[f"{tag}/min"]) and np.array_equal(mx, GOLD[f"{tag}/max"])
        assert np.array_equal(codes, GOLD[f"{tag}/codes"])
    for k, ef in ((10, 32), (5, 50), (10, 10)):
        want = GOLD[f"{tag}/ids_k{k}_ef{ef}"]
        got = ix.batch_search(q, k, ef)
        assert np.mean(np.all(got == want, axis=1)) >= 0.9, (tag, k, ef)
        assert recall(got, want) >= 0.98
    ids, dist = ix.batch_search_with_distance(q, 10, 32)
    assert dist.shape == GOLD[f"{tag}/wd_dist"].shape
    same = np.all(ids == GOLD[f"{tag}/wd_ids"], axis=1)
    assert same.mean() >= 0.9
    if dist.size:
        assert np.allclose(dist[same], GOLD[f"{tag}/wd_dist"][same], rtol=1e-4, atol=1e-6)


@pytest.mark.parametrize("metric,quant", [c for c in CONFIGS if c[0] != "cos"])
def test_insert_remove_match_reference(cabi, metric, quant):
    """GraphUpdateJob semantics incl. quirks: zeros in unfilled slots, SQ insert skips the raw space."""
    tag, ix = _on_golden_graph(cabi, metric, quant)
    q = GOLD["queries"]
    ins = [ix.insert(q[i], 32) for i in range(3)]
    ix.remove(5)
    assert ins == GOLD[f"{tag}/insert_ids"].tolist()
    want_adj = GOLD[f"{tag}/adj_after_update"]
    got_adj = ix.graph()["adj"]
    assert got_adj.shape == want_adj.shape
    assert np.mean(got_adj == want_adj) >= 0.999
    got = ix.batch_search(q, 10, 32)
    assert np.mean(np.all(got == GOLD[f"{tag}/ids_after_update"], axis=1)) >= 0.9
    if quant == "none":
        assert np.allclose(ix.get_data_by_id(600), q[0])
    else:
        with pytest.raises(RuntimeError):
            ix.get_data_by_id(600)  # the raw space was not extended (graph_update_job.hpp:75)


@pytest.mark.parametrize("quant", ["none", "sq8"])
def test_search_after_removes_equals_reference_search_solo_updated(cabi, ref, quant):
    """GraphSearchJob::search_solo_updated (graph_search_job.hpp:373-417) walks removed vertices through the neighbour
    lists saved at removal time. The device keeps a removed vertex's adjacency row frozen at exactly that list, so the
    ordinary traversal must return the reference's ids, step for step."""
    from alayalite_b200.synthetic import integer_grid, low_rank
    if quant == "none":
        base, q = integer_grid(1500, 24, 60, seed=3, hi=10)  # integer coordinates: every fp32 sum exact, ties included
    else:  # code-space sums are not exact: a grid would turn its many exact ties into coin flips
        base, q = low_rank(1500, 24, 60, latent=6, seed=3)
    r = ref.index("l2", quant, 1600, 24)
    r.fit(base.copy(), 40, 1)
    g = r.graph()
    ix = cabi.NativeIndex("l2", quant, 1600)
    ix.fit_spaces(base)
    ix.set_graph(g["adj"], g["levels"], g["lists_flat"], g["ep"])
    # remove hubs that the traversals certainly cross: the most frequent level-0 neighbours
    adj = g["adj"][:1500]
    hubs = np.argsort(-np.bincount(adj[adj != 0xFFFFFFFF], minlength=1500))[:12]
    for v in hubs.tolist():
        r.remove(v)
        ix.remove(v)
    for ef, k in ((32, 10), (64, 32)):
        got, _ = ix.batch_search_with_distance(q, k, ef)  # pool order, no re-rank
        want = np.stack([r.search_updated(q[i], k, ef) for i in range(len(q))])
        if quant == "none":
            assert np.array_equal(got, want), (quant, ef, k)  # integer coordinates: exact, ties included
        else:  # code-space distances are not exact sums: identical up to fp32 summation-order ties
            assert np.mean(np.all(got == want, axis=1)) >= 0.9, (quant, ef, k)
            assert np.mean([len(set(a.tolist()) & set(b.tolist())) / k for a, b in zip(got, want)]) >= 0.98


def test_insert_until_full_and_sdk_errors(cabi):
    from alayalite_b200 import Client

    rng = np.random.default_rng(5)
    base = rng.standard_normal((1000, 16)).astype(np.float32)
    index = Client().create_index("full", capacity=1002)
    index.fit(base, ef_construction=40)
    assert index.insert(rng.standard_normal(16).astype(np.float32)) == 1000  # python/tests/test_update.py:33-61
    assert index.insert(rng.standard_normal(16).astype(np.float32)) == 1001
    with pytest.raises(RuntimeError, match="full"):
        index.insert(rng.standard_normal(16).astype(np.float32))
    with pytest.raises(RuntimeError):
        index.fit(base)  # "An index can be only fitted once"
    with pytest.raises(RuntimeError, match="out of range"):
        index.get_data_by_id(5000)
    with pytest.raises(ValueError):
        index.search(np.zeros(3, np.float32), 1)  # wrong dimension


# ---- file formats -------------------------------------------------------------------------------
@pytest.mark.parametrize("metric,quant", [("l2", "none"), ("ip", "sq8"), ("l2", "sq4")])
def test_files_are_byte_identical_to_the_reference(cabi, ref, tmp_path, metric, quant):
    """Same rows + same graph -> Graph::save / RawSpace::save / SQxSpace::save bytes (Appendix C)."""
    base, q = GOLD["base"], GOLD["queries"]
    r = ref.index(metric, quant, 640, 24)
    r.fit(base.copy(), 40, 1)
    qp = "" if quant == "none" else str(tmp_path / f"ref_{quant}.data")
    r.save(str(tmp_path / "ref.index"), str(tmp_path / "ref_raw.data"), qp)
    tag, ix = _on_golden_graph(cabi, metric, quant)
    oqp = "" if quant == "none" else str(tmp_path / f"our_{quant}.data")
    ix.save(str(tmp_path / "our.index"), str(tmp_path / "our_raw.data"), oqp)
    for a, b in (("ref.index", "our.index"), ("ref_raw.data", "our_raw.data")) + (((f"ref_{quant}.data", f"our_{quant}.data"),) if quant != "none" else ()):
        assert open(tmp_path / a, "rb").read() == open(tmp_path / b, "rb").read(), (a, b)
    # and each side loads the other's files
    r2 = ref.index(metric, quant, 640, 24)
    r2.load(str(tmp_path / "our.index"), str(tmp_path / "our_raw.data"), oqp)
    assert np.array_equal(r2.batch_search(q, 10, 32, 1), r.batch_search(q, 10, 32, 1))
    ix2 = cabi.NativeIndex(metric, quant, 1)  # capacity comes from the file
    ix2.load(str(tmp_path / "ref.index"), str(tmp_path / "ref_raw.data"), qp)
    assert ix2.size == 600 and ix2.dim == 24
    assert np.array_equal(ix2.batch_search(q, 10, 32), ix.batch_search(q, 10, 32))


@pytest.mark.parametrize("metric,quant,index_type", [("ip", "sq8", "nsg"), ("cos", "sq4", "fusion"), ("l2", "none", "hnsw")])
def test_client_save_and_reload(cabi, tmp_path, metric, quant, index_type):
    """python/tests/test_client_with_url.py:77-105: save -> Client(dir) reload -> same answers."""
    from alayalite_b200 import Client
    from alayalite_b200.synthetic import low_rank

    base, q = low_rank(1200, 32, 30, latent=6, seed=9)
    c = Client(str(tmp_path))
    index = c.create_index("idx", metric=metric, quantization_type=quant, index_type=index_type, capacity=1500)
    index.fit(base.copy(), ef_construction=50)
    before = index.batch_search(q, 10, 50)
    c.save_index("idx")
    files = sorted(os.listdir(tmp_path / "idx"))
    assert "schema.json" in files and "raw.data" in files and f"{index_type}_{metric}_32.index" in files
    c2 = Client(str(tmp_path))
    assert c2.list_indices() == ["idx"]
    again = c2.get_index("idx")
    assert again.get_dim() == 32
    assert np.array_equal(again.batch_search(q, 10, 50), before)
    if metric != "cos":
        assert np.allclose(again.get_data_by_id(7), base[7])


# ---- GPU-built graphs: quality vs the reference's own build ------------------------------------------
@pytest.mark.parametrize("metric,quant", [("l2", "none"), ("ip", "none"), ("cos", "none"), ("l2", "sq8"), ("ip", "sq4")])
def test_gpu_build_recall_matches_reference_build(cabi, orc, metric, quant):
    """north_star: recall@10 within 0.005 of the reference's own index at the same search parameters."""
    from alayalite_b200.synthetic import low_rank

    base, q = low_rank(20000, 64, 500, latent=10, seed=17, normalize=(metric != "l2"))
    ix = cabi.NativeIndex(metric, quant, 20000)
    ix.fit(base, 100)
    g = ix.graph()
    deg = (g["adj"] != 0xFFFFFFFF).sum(1)
    assert deg.min() >= 1 and deg.max() <= 32
    assert np.array_equal(g["levels"], orc.hnsw_levels(20000))  # same level sequence and entry point rule
    lv = g["levels"]
    assert g["ep"] == int(np.argmax(lv)) and g["lists_flat"].size == int(lv.sum()) * 32
    o = orc.index(metric, quant, 20000, 64)
    o.fit(base.copy(), 100)  # the reference algorithm, sequential insertion
    gt, _ = cabi.bruteforce_topk("l2" if metric == "l2" else "ip", base, q, 10)
    for ef in (16, 40, 100):
        r_gpu = recall(ix.batch_search(q, 10, ef), gt)
        r_ref = recall(o.batch_search(q, 10, ef, 8), gt)
        assert r_gpu >= r_ref - 0.005, (metric, quant, ef, r_gpu, r_ref)


def test_quickstart_config(cabi):
    """BASELINE.json configs[0]: random 1000 x 128, Index.fit + batch_search top-10 L2;
    python/tests/test_index_types.py:32-40 asserts recall >= 0.9 at the default ef."""
    from alayalite_b200 import Client, calc_gt, calc_recall
    from alayalite_b200.synthetic import uniform

    base, q = uniform(1000, 128, 10, seed=0)
    index = Client().create_index("default")
    index.fit(base)
    res = index.batch_search(q, 10)
    assert res.shape == (10, 10) and res.dtype == np.uint32
    assert calc_recall(res, calc_gt(base, q, 10)) >= 0.9
    assert np.array_equal(index.search(q[0], 10), res[0])


@pytest.mark.parametrize("dtype", [np.int32, np.uint32, np.uint8, np.float64])
def test_other_dtypes(cabi, dtype):
    """python/tests/test_index_types.py:42-80: non-float inputs, recall >= 0.9."""
    from alayalite_b200 import Client, calc_gt, calc_recall

    rng = np.random.default_rng(1)
    base = rng.integers(0, 100, (1000, 64)).astype(dtype)
    q = rng.integers(0, 100, (20, 64)).astype(dtype)
    index = Client().create_index("t", data_type=dtype, id_type=np.uint64)
    index.fit(base)
    res = index.batch_search(q, 10)
    assert res.dtype == np.uint64
    assert calc_recall(res.astype(np.int64), calc_gt(base.astype(np.float32), q.astype(np.float32), 10).astype(np.int64)) >= 0.9
    assert index.get_data_by_id(3).dtype == dtype and np.array_equal(index.get_data_by_id(3), base[3])
