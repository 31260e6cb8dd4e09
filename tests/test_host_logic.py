"""Host-side SDK logic that needs no GPU: parameter validation, schema files, Client registry,
vector-file loaders. Mirrors the reference's python/tests (test_client.py, test_index_types.py
argument checks) -- same calls, same exception types."""

import json
import os

import numpy as np
import pytest

from alayalite_b200 import Client, Index, IndexParams, calc_recall, load_fvecs, load_ivecs
from alayalite_b200.schema import is_collection_url, is_index_url, load_schema, save_schema


def test_index_params_defaults_and_paths(tmp_path):
    p = IndexParams()
    p.fill_none_values()
    assert (p.index_type, p.quantization_type, p.metric, p.capacity, p.max_nbrs) == ("hnsw", "none", "l2", 100000, 32)
    assert p.data_type == np.float32 and p.id_type == np.uint32
    assert p.index_path("d") == os.path.join("d", "hnsw_l2_32.index")
    assert p.data_path("d") == os.path.join("d", "raw.data") and p.quant_path("d") == ""
    q = IndexParams.from_kwargs(index_type="nsg", metric="cosine", quantization_type="sq4", capacity=7, max_nbrs=64, id_type=np.uint64)
    q.fill_none_values()
    assert q.index_path("d").endswith("nsg_cosine_64.index") and q.quant_path("d").endswith("sq4.data")
    d = q.to_json_dict()
    assert d["id_type"] == "uint64" and d["data_type"] == "float32"
    r = IndexParams.from_str_dict(json.loads(json.dumps(d)))
    assert r == q
    native = q.to_cpp_params()
    assert native.capacity_ == 7


@pytest.mark.parametrize("kw", [dict(metric="manhattan"), dict(index_type="ivf"), dict(quantization_type="pq"), dict(capacity=0),
                                dict(max_nbrs=0), dict(max_nbrs=1000), dict(data_type=np.complex64), dict(id_type=np.int16)])
def test_invalid_kwargs_raise_value_error(kw):
    with pytest.raises(ValueError):
        Client().create_index("x", **kw)


def test_client_registry_and_disk_rules(tmp_path):
    c = Client()
    a = c.create_index("a", metric="ip")
    assert c.get_index("a") is a and c.list_indices() == ["a"] and c.get_index("nope") is None
    assert c.get_or_create_index("a") is a and isinstance(c.get_or_create_index("b"), Index)
    with pytest.raises(RuntimeError):
        c.create_index("a")
    with pytest.raises(RuntimeError):
        c.save_index("a")  # no url
    with pytest.raises(RuntimeError):
        c.delete_index("zzz")
    with pytest.raises(RuntimeError):
        c.delete_index("a", delete_on_disk=True)  # removed from the map, but no url for disk ops
    assert c.list_indices() == ["b"]
    c.reset()
    assert c.list_indices() == []
    d = Client(str(tmp_path / "db"))
    with pytest.raises(RuntimeError):
        d.save_index("missing")
    assert os.path.isdir(tmp_path / "db")


def test_index_argument_checks_do_not_need_a_device():
    ix = Index("i", IndexParams())
    with pytest.raises(ValueError):
        ix.search(np.zeros(4, np.float32), 1)  # not fitted
    with pytest.raises(ValueError):
        ix.fit(np.zeros(8, np.float32))  # not 2-D
    jx = Index("j", IndexParams(data_type=np.float32))
    with pytest.raises(ValueError):
        jx.fit(np.zeros((4, 4), np.float64))  # dtype mismatch


def test_schema_files(tmp_path):
    folder = tmp_path / "idx"
    folder.mkdir()
    path = folder / "schema.json"
    assert not is_index_url(str(folder)) and not is_collection_url(str(folder))
    path.write_text(json.dumps({"type": "index", "index": {}}))
    assert is_index_url(str(folder)) and not is_collection_url(str(folder))
    save_schema(str(path), {"type": "collection"})
    assert load_schema(str(path)) == {"type": "collection"} and not os.path.exists(str(path) + ".bak")
    with pytest.raises(FileNotFoundError):
        load_schema(str(folder / "none.json"))
    with pytest.raises(RuntimeError):
        Index.load(str(tmp_path), "does-not-exist")


def test_vecs_loaders_and_recall(tmp_path):
    v = np.arange(12, dtype=np.float32).reshape(3, 4)
    with open(tmp_path / "a.fvecs", "wb") as f:
        for row in v:
            f.write(np.int32(4).tobytes() + row.tobytes())
    assert np.array_equal(load_fvecs(str(tmp_path / "a.fvecs")), v)
    iv = np.arange(6, dtype=np.int32).reshape(2, 3)
    with open(tmp_path / "a.ivecs", "wb") as f:
        for row in iv:
            f.write(np.int32(3).tobytes() + row.tobytes())
    assert np.array_equal(load_ivecs(str(tmp_path / "a.ivecs")), iv)
    res = np.array([[1, 2, 3], [4, 5, 6]])
    assert calc_recall(res, res) == 1.0
    assert calc_recall(res, np.array([[3, 9, 9], [9, 9, 9]])) == pytest.approx(1 / 6)
    assert calc_recall(res, np.array([[7, 8, 9], [7, 8, 9]])) == 0.0
