"""IndexParams and the schema.json helpers.

API and file naming follow /root/reference/python/src/alayalite/schema.py:46-211 so that a folder
saved by either SDK loads in the other: `{index_type}_{metric}_{max_nbrs}.index`, `raw.data`,
`{sq8|sq4}.data`, and `schema.json` = {"type": "index", "index": {...}}.
"""

from __future__ import annotations

import json
import os
import shutil
from dataclasses import dataclass, fields

import numpy as np

from ._alayalitepy import IndexParams as _IndexParams
from . import common as _c

__all__ = ["IndexParams", "load_schema", "save_schema", "is_index_url", "is_collection_url"]

_DEFAULTS = dict(index_type="hnsw", data_type=np.float32, id_type=np.uint32, quantization_type="none", metric="l2",
                 capacity=100000, max_nbrs=32)


@dataclass
class IndexParams:
    """Parameters of one vector index. Unset fields are filled by fill_none_values()."""

    index_type: str = None
    data_type: type = None
    id_type: type = None
    quantization_type: str = None
    metric: str = None
    capacity: int = None
    max_nbrs: int = None

    # ---- file names inside an index folder ----------------------------------------------------
    def index_path(self, folder_uri):
        return os.path.join(folder_uri, f"{self.index_type}_{self.metric}_{self.max_nbrs}.index")

    def data_path(self, folder_uri):
        return os.path.join(folder_uri, "raw.data")

    def quant_path(self, folder_uri):
        return "" if self.quantization_type == "none" else os.path.join(folder_uri, f"{self.quantization_type}.data")

    def fill_none_values(self):
        for f in fields(self):
            if getattr(self, f.name) is None:
                setattr(self, f.name, _DEFAULTS[f.name])

    def to_cpp_params(self):
        # max_nbrs is deliberately not forwarded: the native constructor takes six fields
        # (pybind.cpp:70-81 in the reference) and the graph degree stays 32.
        return _IndexParams(
            index_type_=_c.valid_index_type(self.index_type),
            data_type_=_c.valid_dtype(self.data_type),
            id_type_=_c.valid_id_type(self.id_type),
            quantization_type_=_c.valid_quantization_type(self.quantization_type),
            metric_=_c.valid_metric_type(self.metric),
            capacity_=_c.valid_capacity_type(self.capacity),
        )

    def to_json_dict(self) -> dict:
        return dict(index_type=self.index_type, data_type=np.dtype(self.data_type).name, id_type=np.dtype(self.id_type).name,
                    quantization_type=self.quantization_type, metric=self.metric, capacity=self.capacity, max_nbrs=self.max_nbrs)

    @classmethod
    def from_str_dict(cls, data: dict) -> "IndexParams":
        return cls(index_type=data["index_type"], data_type=np.dtype(data["data_type"]).type, id_type=np.dtype(data["id_type"]).type,
                   quantization_type=data["quantization_type"], metric=data["metric"], capacity=data["capacity"], max_nbrs=data["max_nbrs"])

    @classmethod
    def from_kwargs(cls, **kwargs) -> "IndexParams":
        checks = dict(index_type=_c.valid_index_type, quantization_type=_c.valid_quantization_type, metric=_c.valid_metric_type)
        converts = dict(data_type=_c.valid_dtype, id_type=_c.valid_id_type, capacity=_c.valid_capacity_type, max_nbrs=_c.valid_max_nbrs)
        out = {}
        for f in fields(cls):
            v = kwargs.get(f.name)
            if v is None:
                continue
            if f.name in checks:
                checks[f.name](v)  # validate, keep the caller's spelling
                out[f.name] = v
            else:
                out[f.name] = converts[f.name](v)
        return cls(**out)


def load_schema(url) -> dict:
    if not os.path.exists(url):
        raise FileNotFoundError("The schema file does not exist!")
    with open(url, encoding="utf-8") as f:
        return json.load(f)


def save_schema(schema_url, schema_map):
    backup = schema_url + ".bak"
    shutil.copy2(schema_url, backup)
    with open(schema_url, "w", encoding="utf-8") as f:
        json.dump(schema_map, f, indent=4)
    os.remove(backup)


def _schema_type(url):
    path = os.path.join(url, "schema.json")
    return load_schema(path)["type"] if os.path.exists(path) else None


def is_index_url(url):
    return _schema_type(url) == "index"


def is_collection_url(url):
    return _schema_type(url) == "collection"
