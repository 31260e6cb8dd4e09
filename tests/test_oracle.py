"""CPU tests of the oracle (oracle/restate.cpp): pinned against (1) the literal known answers of the
reference's own unit tests, (2) golden outputs of the compiled reference (tests/golden/ref_small.npz,
made by tests/golden/make_golden.py) and (3) the compiled reference itself where oracle/_ref exists."""

import json
import os

import numpy as np
import pytest

HERE = os.path.dirname(os.path.abspath(__file__))
KA = json.load(open(os.path.join(HERE, "golden", "known_answers.json")))
GOLD = np.load(os.path.join(HERE, "golden", "ref_small.npz"))
CONFIGS = [("l2", "none"), ("ip", "none"), ("cos", "none"), ("l2", "sq8"), ("ip", "sq8"), ("l2", "sq4"), ("cos", "sq4")]


def checkers(orc, request):
    out = [orc]
    try:
        out.append(request.getfixturevalue("ref"))
    except pytest.skip.Exception:
        pass
    return out


# ---- (1) known answers --------------------------------------------------------------------------
def test_known_raw_distances(orc, request):
    for chk in checkers(orc, request):
        for case in KA["raw_l2"]:
            assert chk.dist("l2_sqr", np.float32(case["x"]), np.float32(case["y"])) == case["expected"]


def test_known_sq_space_distance(orc, request):
    data = np.array(KA["sq_space_distance"]["data"], np.float32)
    for chk in checkers(orc, request):
        for bits in (8, 4):
            mn, mx, codes = chk.fit_encode(bits, data)
            assert chk.dist(f"l2_sqr_sq{bits}", codes[0], codes[1], mn, mx) == pytest.approx(KA["sq_space_distance"]["expected"], rel=1e-6)


def test_known_quantizer(orc, request):
    fit, qz, enc = KA["quantizer_fit"], KA["quantize"], KA["encode"]
    for chk in checkers(orc, request):
        for bits in (8, 4):
            mn, mx, _ = chk.fit_encode(bits, np.array(fit["data"], np.float32))
            assert mn.tolist() == fit["min"] and mx.tolist() == fit["max"]
            for c in qz["cases"]:
                assert chk.quantize(bits, c["v"], qz["min"], qz["max"]) == c[f"sq{bits}"]
            # encode with pinned min/max: fit on rows that realise exactly that range
            rows = np.array([enc["min"], enc["max"], enc["raw"]], np.float32)
            _, _, codes = chk.fit_encode(bits, rows)
            assert codes[2].tolist() == enc[f"sq{bits}"]


def test_known_sq_extremes(orc, request):
    e = KA["sq_extremes"]
    d = e["dim"]
    mn, mx = np.full(d, e["min"], np.float32), np.full(d, e["max"], np.float32)
    for chk in checkers(orc, request):
        assert abs(chk.dist("l2_sqr_sq8", np.zeros(d, np.uint8), np.full(d, 255, np.uint8), mn, mx) - e["expected_l2"]) < e["tolerance"]
        assert abs(chk.dist("l2_sqr_sq4", np.zeros(d // 2, np.uint8), np.full(d // 2, 255, np.uint8), mn, mx) - e["expected_l2"]) < e["tolerance"]


def test_known_linear_pool(orc, request):
    for chk in checkers(orc, request):
        for case in KA["linear_pool"]:
            ops = [tuple(o) for o in case["ops"]]
            popped, ids, _ = chk.linear_pool_replay(100, case["capacity"], ops)
            assert popped.tolist() == case["popped"], case["source"]
            assert len(ids) == case["final_size"] and ids.tolist() == case["final_ids"], case["source"]


def test_known_find_exact_gt(orc, request):
    c = KA["find_exact_gt"]
    for chk in checkers(orc, request):
        got = chk.find_exact_gt(np.float32(c["queries"]), np.float32(c["data"]), c["topk"])
        assert got.tolist() == c["expected"]


def test_known_sq4_swapped_nibbles(orc, request):
    c = KA["sq4_swapped_nibbles"]
    for chk in checkers(orc, request):
        mn, mx, codes = chk.fit_encode(4, np.float32(c["data"]))
        i, j = c["pair"]
        assert chk.dist("l2_sqr_sq4", codes[i], codes[j], mn, mx) == pytest.approx(c["expected"], rel=1e-5)


# ---- (2) golden outputs of the compiled reference --------------------------------------------------
@pytest.mark.parametrize("metric,quant", CONFIGS)
def test_oracle_reproduces_golden(orc, metric, quant):
    tag = f"{metric}_{quant}"
    base, q = GOLD["base"], GOLD["queries"]
    ix = orc.index(metric, quant, 640, 24)
    ix.fit(base.copy(), 40)
    g = ix.graph()
    cos = metric == "cos"  # the reference's normalize() is summation-order dependent under -Ofast
    if not cos:
        assert np.array_equal(g["adj"], GOLD[f"{tag}/adj"])
        assert np.array_equal(g["levels"], GOLD[f"{tag}/levels"]) and np.array_equal(g["lists_flat"], GOLD[f"{tag}/lists"])
        assert np.array_equal(ix.raw_rows(), GOLD[f"{tag}/raw"])
    else:
        assert np.allclose(ix.raw_rows(), GOLD[f"{tag}/raw"], rtol=1e-6, atol=1e-7)
        assert np.mean(g["adj"] == GOLD[f"{tag}/adj"]) > 0.98
    assert g["ep"] == int(GOLD[f"{tag}/ep"])
    if quant != "none":
        mn, mx = ix.minmax()
        if not cos:
            assert np.array_equal(mn, GOLD[f"{tag}/min"]) and np.array_equal(mx, GOLD[f"{tag}/max"])
            assert np.array_equal(ix.codes(), GOLD[f"{tag}/codes"])
        else:
            assert np.allclose(mn, GOLD[f"{tag}/min"], rtol=1e-5) and np.mean(ix.codes() == GOLD[f"{tag}/codes"]) > 0.99
    for k, ef in ((10, 32), (5, 50), (10, 10)):
        want = GOLD[f"{tag}/ids_k{k}_ef{ef}"]
        got = ix.batch_search(q, k, ef)
        if cos:
            assert np.mean(np.all(got == want, axis=1)) >= 0.9
        else:
            assert np.array_equal(got, want), (tag, k, ef)
    ids, dist = ix.batch_search_with_distance(q, 10, 32)
    if not cos:
        assert np.array_equal(ids, GOLD[f"{tag}/wd_ids"])
    assert dist.shape == GOLD[f"{tag}/wd_dist"].shape
    if dist.size and not cos:
        assert np.allclose(dist, GOLD[f"{tag}/wd_dist"], rtol=1e-5, atol=1e-6)
    pairs = GOLD[f"{tag}/pairs"]
    got = np.array([ix.space_distance(int(i), int(j)) for i, j in pairs], np.float32)
    assert np.allclose(got, GOLD[f"{tag}/pair_dist"], rtol=1e-4, atol=1e-5)


@pytest.mark.parametrize("metric,quant", [c for c in CONFIGS if c[0] != "cos"])
def test_oracle_update_reproduces_golden(orc, metric, quant):
    """GraphUpdateJob::insert_and_update / remove (graph_update_job.hpp:66-137)."""
    tag = f"{metric}_{quant}"
    base, q = GOLD["base"], GOLD["queries"]
    ix = orc.index(metric, quant, 640, 24)
    ix.fit(base.copy(), 40)
    ins = [ix.insert(q[i], 32) for i in range(3)]
    ix.remove(5)
    assert ins == GOLD[f"{tag}/insert_ids"].tolist()
    assert np.array_equal(ix.graph()["adj"], GOLD[f"{tag}/adj_after_update"])
    assert np.array_equal(ix.batch_search(q, 10, 32), GOLD[f"{tag}/ids_after_update"])


def test_oracle_gt_reproduces_golden(orc):
    assert np.array_equal(orc.find_exact_gt(GOLD["queries"], GOLD["base"], 10), GOLD["gt_k10"])


def test_hnsw_level_sequence(orc):
    """hnswlib.hpp:182-186 with seed 100: the golden graphs were built with this sequence."""
    lv = orc.hnsw_levels(600)
    assert np.array_equal(lv, GOLD["l2_none/levels"])
    frac = np.mean(orc.hnsw_levels(200000) > 0)
    assert abs(frac - 1 / 16) < 0.003


# ---- (3) live pin against the compiled reference ------------------------------------------------------
@pytest.mark.parametrize("dim", [1, 3, 7, 15, 16, 17, 31, 33, 63, 65, 127, 128, 129, 960])
def test_distances_vs_ref(orc, ref, dim):
    rng = np.random.default_rng(dim)
    x, y = rng.standard_normal(dim).astype(np.float32), rng.standard_normal(dim).astype(np.float32)
    for name in ("l2_sqr", "ip_sqr"):
        a = orc.dist(name, x, y)
        assert a == pytest.approx(ref.dist(name, x, y), rel=1e-5, abs=1e-5)
        assert a == pytest.approx(ref.dist(name, x, y, generic=True), rel=1e-5, abs=1e-5)
    mn = rng.uniform(-1, 0, dim).astype(np.float32)
    mx = (mn + rng.uniform(0.1, 2, dim)).astype(np.float32)
    for bits in (8, 4):
        rb = dim if bits == 8 else (dim + 1) // 2
        cx, cy = rng.integers(0, 256, rb).astype(np.uint8), rng.integers(0, 256, rb).astype(np.uint8)
        if bits == 4 and dim % 2:
            cx[-1] &= 0xF0
            cy[-1] &= 0xF0
        for name in (f"l2_sqr_sq{bits}", f"ip_sqr_sq{bits}"):
            assert orc.dist(name, cx, cy, mn, mx) == pytest.approx(ref.dist(name, cx, cy, mn, mx), rel=2e-5, abs=1e-4)


@pytest.mark.parametrize("metric,quant", [("l2", "none"), ("ip", "sq8"), ("l2", "sq4")])
def test_build_and_search_vs_ref(orc, ref, metric, quant):
    from alayalite_b200.synthetic import low_rank

    base, q = low_rank(2500, 40, 60, latent=8, seed=21)
    a, b = orc.index(metric, quant, 3000, 40), ref.index(metric, quant, 3000, 40)
    a.fit(base.copy(), 80)
    b.fit(base.copy(), 80, 1)
    ga, gb = a.graph(), b.graph()
    assert np.array_equal(ga["adj"], gb["adj"]) and np.array_equal(ga["lists_flat"], gb["lists_flat"]) and ga["ep"] == gb["ep"]
    for k, ef in ((10, 16), (10, 100), (50, 50)):
        want = b.batch_search(q, k, ef, 3)
        assert np.array_equal(a.batch_search(q, k, ef, 2), want)
        assert np.array_equal(b.batch_search(q, k, ef, 2, solo=True), want)  # coroutine path == search_solo
    ia, da, sa = a.traverse(q[0], 64)
    ib, db, sb = b.traverse(q[0], 64)
    assert np.array_equal(ia, ib) and sa == sb and np.allclose(da, db, rtol=1e-5)
