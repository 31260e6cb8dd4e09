"""Argument validation shared by the SDK classes.

Same accepted values and the same ValueError behaviour as the reference SDK
(/root/reference/python/src/alayalite/common.py:80-187); the native enums come from this package's
own pybind module.
"""

from __future__ import annotations

import numpy as np

from ._alayalitepy import IndexType as _IndexType
from ._alayalitepy import MetricType as _MetricType
from ._alayalitepy import QuantizationType as _QuantizationType

ID_DTYPES = (np.uint64, np.uint32)
VECTOR_DTYPES = (np.float32, np.int8, np.uint8, np.float64, np.int32, np.uint32)
METRICS = {"l2": _MetricType.L2, "euclidean": _MetricType.L2, "ip": _MetricType.IP, "cos": _MetricType.COS, "cosine": _MetricType.COS}
INDEX_TYPES = {"hnsw": _IndexType.HNSW, "nsg": _IndexType.NSG, "fusion": _IndexType.FUSION}
QUANTIZATIONS = {None: _QuantizationType.NONE, "none": _QuantizationType.NONE, "sq8": _QuantizationType.SQ8,
                 "sq4": _QuantizationType.SQ4, "rabitq": _QuantizationType.RABITQ}

__all__ = ["valid_dtype", "valid_id_type", "valid_metric_type", "valid_index_type", "valid_capacity_type",
           "valid_quantization_type", "valid_max_nbrs", "require"]


def require(condition: bool, message: str) -> None:
    """The SDK reports bad arguments as ValueError (common.py:185-187 in the reference)."""
    if not condition:
        raise ValueError(message)


def _castable(dtype, allowed) -> bool:
    return any(np.can_cast(dtype, a) for a in allowed)


def valid_dtype(dtype) -> np.dtype:
    require(_castable(dtype, VECTOR_DTYPES), "Vector dtype must be one of float32, int8, uint8, float64, int32, uint32")
    return np.dtype(dtype)


def valid_id_type(id_type) -> np.dtype:
    require(_castable(id_type, ID_DTYPES), "ID dtype must be one of uint64, uint32")
    return np.dtype(id_type)


def valid_capacity_type(capacity):
    require(capacity > 0, "Capacity must be greater than 0")
    return capacity


def valid_max_nbrs(max_nbrs):
    require(0 < max_nbrs < 1000, "Max neighbors must be greater than 0 and less than 1000")
    return max_nbrs


def _lookup(table, key, what):
    k = key.lower() if isinstance(key, str) else key
    require(k in table, f"{what} must be one of {[x for x in table if x is not None]}")
    return table[k]


def valid_metric_type(metric: str):
    return _lookup(METRICS, metric, "Distance metric")


def valid_index_type(index: str):
    return _lookup(INDEX_TYPES, index, "Index type")


def valid_quantization_type(quantization):
    return _lookup(QUANTIZATIONS, quantization, "Quantization type")
