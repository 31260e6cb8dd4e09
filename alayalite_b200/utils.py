"""Ground truth, recall and vector-file helpers.

calc_gt keeps the reference's contract (/root/reference/python/src/alayalite/utils.py:99-105:
exact L2 top-k ids as int32 [nq, topk]) but runs the exhaustive search on the GPU through
ab_bruteforce_topk; ids are ordered by (distance, id) where the reference leaves ties unspecified.
"""

from __future__ import annotations

import numpy as np

from . import _cabi

__all__ = ["load_fvecs", "load_ivecs", "calc_recall", "calc_gt"]


def _load_vecs(path, dtype):
    raw = np.fromfile(path, dtype=np.int32)
    if raw.size == 0:
        return np.zeros((0, 0), dtype)
    dim = int(raw[0])
    return raw.reshape(-1, dim + 1)[:, 1:].copy().view(dtype)


def load_fvecs(file_path):
    """<int32 dim><dim x float32> records -> [n, dim] float32."""
    return _load_vecs(file_path, np.float32)


def load_ivecs(file_path):
    """<int32 dim><dim x int32> records -> [n, dim] int32."""
    return _load_vecs(file_path, np.int32)


def calc_recall(result, gt_data):
    """Mean over queries of |result_i intersect gt_i| / topk (utils.py:78-84)."""
    rows, cols = result.shape
    hits = sum(len(set(result[i]) & set(gt_data[i])) for i in range(rows))
    return 1.0 * hits / (rows * cols)


def calc_gt(data, query, topk):
    ids, _ = _cabi.bruteforce_topk("l2", data, query, topk)
    return ids.astype(np.int32)
