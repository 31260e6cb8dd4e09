"""Index: the user-facing handle of one vector index.

Method names, argument meaning, defaults and error behaviour follow the reference SDK
(/root/reference/python/src/alayalite/index.py:35-231); the work is done by
_alayalitepy.PyIndexInterface, i.e. by libalaya_b200.so on the GPU.
"""

from __future__ import annotations

import os

import numpy as np

from ._alayalitepy import PyIndexInterface as _PyIndexInterface
from .common import require
from .schema import IndexParams, load_schema

_FULL = (0xFFFFFFFF, 0xFFFFFFFFFFFFFFFF, -1)


class Index:
    def __init__(self, name: str = "default", params: IndexParams = None):
        self._name = name
        self._params = params if params is not None else IndexParams()
        self._native = None
        self._dim = None

    def get_params(self) -> IndexParams:
        return self._params

    def get_dim(self):
        return self._dim

    def get_dtype(self):
        return self._params.data_type

    def _need_native(self):
        require(self._native is not None, "Index is not init yet")

    def _need_dim(self, got, what):
        require(got == self._dim,
                f"{what} dimension must match the dimension of the vectors used to fit the index."
                f"fit data dimension: {self._dim}, {what} dimension: {got}")

    def fit(self, vectors, ef_construction: int = 100, num_threads: int = 1):
        """Build the index from a 2-D array. An index can be fitted once."""
        if self._native is not None:
            raise RuntimeError("An index can be only fitted once")
        require(vectors.ndim == 2, "vectors must be a 2D array")
        dtype = np.array(vectors).dtype
        if self._params.data_type is None:
            self._params.data_type = dtype
        elif self._params.data_type != dtype:
            raise ValueError(f"Data type mismatch: {self._params.data_type} vs {dtype}")
        self._params.fill_none_values()
        self._dim = vectors.shape[1]
        self._native = _PyIndexInterface(self._params.to_cpp_params())
        self._native.fit(vectors, ef_construction, num_threads)

    def get_data_by_id(self, vector_id: int):
        return self._native.get_data_by_id(vector_id)

    def insert(self, vectors, ef: int = 100):
        self._need_native()
        require(vectors.ndim == 1, "vectors must be a 1D array")
        self._need_dim(vectors.shape[0], "vectors")
        new_id = self._native.insert(vectors, ef)
        if new_id in _FULL:
            raise RuntimeError("The index is full, cannot insert more vectors")
        return new_id

    def remove(self, vector_id: int) -> None:
        self._need_native()
        self._native.remove(vector_id)

    def search(self, query, topk: int, ef_search: int = 100):
        self._need_native()
        require(query.ndim == 1, "query must be a 1D array")
        self._need_dim(query.shape[0], "query")
        return self._native.search(query, topk, ef_search)

    def batch_search(self, queries, topk: int, ef_search: int = 100, num_threads: int = 1):
        self._need_native()
        require(queries.ndim == 2, "queries must be a 2D array")
        self._need_dim(queries.shape[1], "query")
        return self._native.batch_search(queries, topk, ef_search, num_threads)

    def batch_search_with_distance(self, queries, topk: int, ef_search: int = 100, num_threads: int = 1):
        self._need_native()
        require(queries.ndim == 2, "queries must be a 2D array")
        self._need_dim(queries.shape[1], "query")
        return self._native.batch_search_with_distance(queries, topk, ef_search, num_threads)

    def save(self, url) -> dict:
        os.makedirs(url, exist_ok=True)
        p = self._params
        self._native.save(p.index_path(url), p.data_path(url), p.quant_path(url))
        return {"type": "index", "index": p.to_json_dict()}

    @classmethod
    def load(cls, url, name):
        folder = os.path.join(url, name)
        if not os.path.exists(folder):
            raise RuntimeError("The index file does not exist")
        params = IndexParams.from_str_dict(load_schema(os.path.join(folder, "schema.json"))["index"])
        self = cls(name, params)
        self._native = _PyIndexInterface(params.to_cpp_params())
        self._native.load(params.index_path(folder), params.data_path(folder), params.quant_path(folder))
        self._dim = self._native.get_data_dim()
        return self
