"""Collection: documents + metadata keyed by caller ids on top of a device Index.

Same public surface and return shapes as the reference's Collection
(/root/reference/python/src/alayalite/collection.py:33-282) -- batch_query / filter_query / insert / upsert /
delete_by_id / get_by_id / delete_by_filter / reindex / save / load / set_metric / get_index_params -- so that
Client.create_collection users can switch. The vector work is Index.fit / insert /
remove / batch_search_with_distance, i.e. the GPU path; the record store here is an insertion-ordered dict
(no DataFrame scans on the query path). `collection.pkl` keeps the reference's layout (a pandas DataFrame with
columns id / document / metadata plus the two id maps), so either implementation can load the other's files.
"""

from __future__ import annotations

import os
import pickle
from typing import List, Optional

import numpy as np

from .common import require as _assert
from .index import Index
from .schema import IndexParams, load_schema

__all__ = ["Collection"]

_COLUMNS = ("id", "document", "metadata")


class Collection:
    def __init__(self, name: str, index_params: IndexParams = None):
        self._name = name
        self._params = index_params if index_params is not None else IndexParams()
        self._index = None
        self._records = {}   # item id -> (document, metadata), insertion ordered
        self._inner_of = {}  # item id -> row id in the index
        self._outer_of = {}  # row id in the index -> item id

    # -- helpers --------------------------------------------------------------------------------
    def _columns(self, ids) -> dict:
        out = {c: [] for c in _COLUMNS}
        for item_id in ids:
            document, metadata = self._records[item_id]
            out["id"].append(item_id)
            out["document"].append(document)
            out["metadata"].append(metadata)
        return out

    @staticmethod
    def _matches(metadata, metadata_filter) -> bool:
        return all(metadata.get(k) == v for k, v in metadata_filter.items())

    def _bind(self, item_id, row_id):
        self._inner_of[item_id] = row_id
        self._outer_of[row_id] = item_id

    # -- queries --------------------------------------------------------------------------------
    def batch_query(self, vectors: List[List[float]], limit: int, ef_search: int = 100, num_threads: int = 1) -> dict:
        """Nearest items of every query vector: {"id": [[...]], "document": ..., "metadata": ..., "distance": ...}."""
        _assert(self._index is not None, "Index is not initialized yet")
        _assert(len(vectors) > 0, "vectors must not be empty")
        _assert(len(vectors[0]) == self._index.get_dim(), "Vector dimension must match the index dimension.")
        _assert(num_threads > 0, "num_threads must be greater than 0")
        _assert(ef_search >= limit, "ef_search must be greater than or equal to limit")
        rows, dists = self._index.batch_search_with_distance(np.asarray(vectors, dtype=np.float32), limit, ef_search,
                                                             num_threads)
        out = {"id": [], "document": [], "metadata": [], "distance": []}
        for q, row_ids in enumerate(rows):
            hits = [self._outer_of[int(r)] for r in row_ids if int(r) in self._outer_of]
            cols = self._columns(hits)
            for c in _COLUMNS:
                out[c].append(cols[c])
            # the reference returns the whole distance row whenever at least one hit survives the id map
            out["distance"].append(np.asarray(dists[q]).tolist() if hits and len(dists) else [])
        return out

    def filter_query(self, metadata_filter: dict, limit: Optional[int] = None) -> dict:
        ids = [i for i, (_, md) in self._records.items() if self._matches(md, metadata_filter)]
        return self._columns(ids if limit is None else ids[:limit])

    def get_by_id(self, ids: List[str]) -> dict:
        wanted = set(ids or [])
        return self._columns([i for i in self._records if i in wanted])

    # -- updates --------------------------------------------------------------------------------
    def insert(self, items: List[tuple]):
        """items: (id, document, embedding, metadata). The first call fits the index on the whole batch."""
        if not items:
            return
        if self._index is None:
            self._params.data_type = np.array(items[0][2]).dtype  # as the reference: avoid a float64 default
            self._index = Index(self._name, self._params)
            self._index.fit(np.array([it[2] for it in items]), ef_construction=100, num_threads=1)
            for row_id, (item_id, document, _, metadata) in enumerate(items):
                self._records[item_id] = (document, metadata)
                self._bind(item_id, row_id)
            return
        dtype = self._index.get_dtype()
        for item_id, document, embedding, metadata in items:
            self._records[item_id] = (document, metadata)
            self._bind(item_id, self._index.insert(np.array(embedding, dtype=dtype)))

    def upsert(self, items: List[tuple]):
        if not items:
            return
        if self._index is None:
            self.insert(items)
            return
        fresh = []
        dtype = self._index.get_dtype()
        for item_id, document, embedding, metadata in items:
            if item_id not in self._inner_of:
                fresh.append((item_id, document, embedding, metadata))
                continue
            # the reference leaves the tombstoned row mapped to the item (collection.py:139-143), which
            # resurrects the old vector on reindex(); the stale mapping is dropped here
            self._index.remove(self._inner_of[item_id])
            del self._outer_of[self._inner_of[item_id]]
            self._bind(item_id, self._index.insert(np.array(embedding, dtype=dtype)))
            self._records[item_id] = (document, metadata)
        self.insert(fresh)

    def delete_by_id(self, ids: List[str]):
        for item_id in ids or []:
            self._records.pop(item_id, None)
            if item_id in self._inner_of:
                row_id = self._inner_of.pop(item_id)
                self._index.remove(row_id)
                del self._outer_of[row_id]

    def delete_by_filter(self, metadata_filter: dict):
        self.delete_by_id([i for i, (_, md) in self._records.items() if self._matches(md, metadata_filter)])

    def reindex(self):
        """Rebuild the index from the live rows (drops tombstones) and renumber the row ids."""
        params = self._index.get_params()
        old_rows = list(self._outer_of.keys())
        vectors = np.array([self._index.get_data_by_id(r) for r in old_rows])
        self._index = Index(self._name, params)
        self._index.fit(vectors)
        outer = {}
        for new_row, old_row in enumerate(old_rows):
            item_id = self._outer_of[old_row]
            self._inner_of[item_id] = new_row
            outer[new_row] = item_id
        self._outer_of = outer

    # -- persistence ----------------------------------------------------------------------------
    def save(self, url) -> dict:
        import pandas as pd

        os.makedirs(url, exist_ok=True)
        frame = pd.DataFrame(self._columns(list(self._records)), columns=list(_COLUMNS))
        with open(os.path.join(url, "collection.pkl"), "wb") as f:
            pickle.dump({"dataframe": frame, "outer_inner_map": self._inner_of, "inner_outer_map": self._outer_of}, f)
        schema = self._index.save(url)
        schema["type"] = "collection"
        return schema

    @classmethod
    def load(cls, url, name):
        folder = os.path.join(url, name)
        if not os.path.exists(folder):
            raise RuntimeError(f"Collection {name} does not exist")
        if load_schema(os.path.join(folder, "schema.json")).get("type") != "collection":
            raise RuntimeError(f"{name} is not a collection")
        self = cls(name)
        with open(os.path.join(folder, "collection.pkl"), "rb") as f:
            blob = pickle.load(f)
        frame = blob["dataframe"]
        for item_id, document, metadata in zip(frame["id"], frame["document"], frame["metadata"]):
            self._records[item_id] = (document, metadata)
        self._inner_of = dict(blob["outer_inner_map"])
        self._outer_of = dict(blob["inner_outer_map"])
        self._index = Index.load(url, name)
        self._params = self._index.get_params()
        return self

    def set_metric(self, metric: str):
        if self._index is not None:
            raise RuntimeError("Cannot change metric after index is created")
        self._params.metric = metric

    def get_index_params(self):
        return self._params
