"""Collection / Client collections. The record store is checked on the CPU with a stand-in
index; the GPU tests run the scenarios of the reference's python/tests/test_collection.py through the real Index."""
import os
import pickle

import numpy as np
import pytest

from alayalite_b200 import collection as collection_mod
from alayalite_b200.collection import Collection
from alayalite_b200.schema import IndexParams


class _ExactIndex:
    """Test double for Index: exact L2 search in numpy, tombstones, same method set as Index."""
    saved = {}

    def __init__(self, name="default", params=None):
        self.name, self.params, self.rows, self.dead = name, params or IndexParams(), None, set()

    def get_params(self): return self.params
    def get_dtype(self): return self.params.data_type
    def get_dim(self): return self.rows.shape[1]
    def fit(self, vectors, ef_construction=100, num_threads=1): self.rows = np.asarray(vectors, dtype=np.float32).copy()
    def get_data_by_id(self, i): return self.rows[i]

    def insert(self, v, ef=100):
        self.rows = np.vstack([self.rows, np.asarray(v, dtype=np.float32)[None]])
        return len(self.rows) - 1

    def remove(self, i): self.dead.add(int(i))

    def batch_search_with_distance(self, q, topk, ef, threads):
        d = ((q[:, None, :] - self.rows[None]) ** 2).sum(-1)
        d[:, sorted(self.dead)] = np.finfo(np.float32).max
        ids = np.argsort(d, axis=1, kind="stable")[:, :topk]
        return ids.astype(np.uint32), np.take_along_axis(d, ids, axis=1)

    def save(self, url):
        _ExactIndex.saved[url] = self
        return {"type": "index", "index": {}}

    @classmethod
    def load(cls, url, name): return cls.saved[os.path.join(url, name)]


@pytest.fixture
def coll(monkeypatch):
    monkeypatch.setattr(collection_mod, "Index", _ExactIndex)
    c = Collection("docs")
    c.insert([(f"k{i}", f"doc {i}", [float(i), 0.0, 1.0], {"parity": i % 2, "i": i}) for i in range(8)])
    return c


def test_store_insert_filter_get(coll):
    assert coll.filter_query({})["id"] == [f"k{i}" for i in range(8)]
    assert coll.filter_query({"parity": 1}, limit=2)["document"] == ["doc 1", "doc 3"]
    got = coll.get_by_id(["k5", "k2", "nope"])
    assert got == {"id": ["k2", "k5"], "document": ["doc 2", "doc 5"], "metadata": [{"parity": 0, "i": 2}, {"parity": 1, "i": 5}]}
    assert coll.get_by_id([]) == {"id": [], "document": [], "metadata": []}
    coll.insert([])
    coll.upsert([])


def test_store_batch_query_orders_by_distance_and_skips_unmapped(coll):
    out = coll.batch_query([[2.2, 0.0, 1.0], [6.9, 0.0, 1.0]], limit=3, ef_search=10)
    assert out["id"] == [["k2", "k3", "k1"], ["k7", "k6", "k5"]]
    assert out["document"][0] == ["doc 2", "doc 3", "doc 1"]
    assert np.allclose(out["distance"][0], [0.04, 0.64, 1.44], atol=1e-5)
    coll.delete_by_id(["k2", "k3"])
    out = coll.batch_query([[2.2, 0.0, 1.0]], limit=2, ef_search=10)
    assert out["id"] == [["k1", "k4"]]
    with pytest.raises(ValueError):
        coll.batch_query([[1.0, 2.0]], limit=1)
    with pytest.raises(ValueError):
        coll.batch_query([[1.0, 2.0, 3.0]], limit=5, ef_search=2)


def test_store_upsert_delete_reindex(coll):
    coll.upsert([("k1", "new 1", [100.0, 0.0, 1.0], {"parity": 9}), ("k8", "doc 8", [8.0, 0.0, 1.0], {"parity": 0})])
    assert coll.get_by_id(["k1"])["document"] == ["new 1"]
    assert coll.batch_query([[100.0, 0.0, 1.0]], limit=1, ef_search=10)["id"] == [["k1"]]
    assert coll.batch_query([[1.0, 0.0, 1.0]], limit=1, ef_search=10)["id"][0][0] != "k1"   # old row is tombstoned
    coll.delete_by_filter({"parity": 0})
    assert coll.filter_query({})["id"] == ["k1", "k3", "k5", "k7"]
    coll.reindex()
    out = coll.batch_query([[3.1, 0.0, 1.0]], limit=4, ef_search=10)
    assert out["id"] == [["k3", "k5", "k7", "k1"]]
    with pytest.raises(RuntimeError):
        coll.set_metric("ip")
    fresh = Collection("x")
    fresh.set_metric("ip")
    assert fresh.get_index_params().metric == "ip"


def test_store_save_load_layout(coll, tmp_path):
    folder = tmp_path / "docs"
    schema = coll.save(str(folder))
    assert schema["type"] == "collection"
    blob = pickle.load(open(folder / "collection.pkl", "rb"))
    assert list(blob["dataframe"].columns) == ["id", "document", "metadata"]      # the reference's pickle layout
    assert blob["outer_inner_map"]["k3"] == 3 and blob["inner_outer_map"][3] == "k3"
    import json
    json.dump(schema, open(folder / "schema.json", "w"))
    back = Collection.load(str(tmp_path), "docs")
    assert back.filter_query({"parity": 1})["id"] == ["k1", "k3", "k5", "k7"]
    assert back.batch_query([[4.1, 0.0, 1.0]], limit=1, ef_search=10)["id"] == [["k4"]]
    with pytest.raises(RuntimeError):
        Collection.load(str(tmp_path), "missing")


# ---- through the real Index ---------------------------------------------------------------------
@pytest.mark.gpu
def test_gpu_tiny_collections():
    c = Collection("tiny")
    c.insert([(1, "Document 1", np.array([0.1, 0.2, 0.3]), {}), (2, "Document 2", np.array([0.4, 0.5, 0.6]), {})])
    out = c.batch_query([[0.1, 0.2, 0.3]], limit=1, ef_search=10)
    assert len(out["id"]) == 1 and out["id"][0][0] == 1
    c = Collection("three")
    c.insert([(1, "d1", [1.0, 2.0, 3.0], {}), (2, "d2", [4.0, 5.0, 6.0], {"category": "B"}), (3, "d3", [12.0, 32.0, 31.0], {"category": "C"})])
    assert [int(x) for x in c.batch_query([[12.0, 32.0, 31.0]], limit=1, ef_search=10)["id"][0]] == [3]
    multi = c.batch_query([[12.0, 32.0, 31.0], [4.0, 5.0, 6.0]], limit=3, ef_search=10)
    assert [int(x) for x in multi["id"][0]] == [3, 2, 1] and [int(x) for x in multi["id"][1]] == [2, 1, 3]
    c = Collection("upd")
    c.insert([(1, "Old Doc", np.array([0.1, 0.2, 0.3]), {})])
    c.upsert([(1, "New Doc", np.array([0.2, 0.3, 0.4]), {})])
    assert c.get_by_id([1])["document"] == ["New Doc"]
    c.delete_by_id([1])
    assert c.get_by_id([1])["id"] == []


@pytest.mark.gpu
def test_gpu_mixed_operations_and_reindex(tmp_path):
    from alayalite_b200 import Client, calc_gt, calc_recall
    rng = np.random.default_rng(3)
    emb = rng.random((999, 100))
    uuids = [f"id_{i}" for i in range(1, 1000)]
    docs = [f"Document {i}" for i in range(1, 1000)]
    client = Client(str(tmp_path))
    c = client.create_collection("mixed")
    with pytest.raises(RuntimeError):
        client.create_index("mixed")
    c.insert(list(zip(uuids, docs, emb.tolist(), [{} for _ in uuids])))
    gone = uuids[10:500:2]
    c.delete_by_id(gone)
    assert c.get_by_id(gone)["id"] == []
    out = c.batch_query([emb[5].tolist()], limit=10, ef_search=50)
    assert out["id"][0][0] == uuids[5] and out["document"][0][0] == docs[5] and abs(out["distance"][0][0]) < 1e-5
    assert not set(out["id"][0]) & set(gone)
    client.save_collection("mixed")
    again = Client(str(tmp_path))
    assert again.list_collections() == ["mixed"] and again.list_indices() == []
    out2 = again.get_collection("mixed").batch_query([emb[5].tolist()], limit=10, ef_search=50)
    assert out2["id"] == out["id"]

    total, dim = 1000, 128
    vec = rng.random((total, dim)).astype(np.float32)
    big = Collection("big")
    big.insert(list(zip(range(total), [f"Doc {i}" for i in range(total)], vec, [{} for _ in range(total)])))
    q = vec[rng.choice(total, 20, replace=False)]
    ids = np.array(big.batch_query(q.tolist(), limit=10, ef_search=100)["id"], dtype=int)
    assert calc_recall(ids, calc_gt(vec, q, 10)) >= 0.9
    big.delete_by_id(list(range(900)))
    big.reindex()
    q = vec[900 + rng.choice(100, 10, replace=False)]
    ids = np.array(big.batch_query(q.tolist(), limit=10, ef_search=100)["id"], dtype=int)
    assert ids.min() >= 900
    assert calc_recall(ids - 900, calc_gt(vec[900:], q, 10)) >= 0.9
