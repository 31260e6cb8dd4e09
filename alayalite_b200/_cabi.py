"""ctypes view of the C ABI (include/alaya_b200.h) for tools that sit below the SDK: the parity
tests, bench.py and the multi-GPU sharding helper. Host numpy arrays in, numpy arrays out; the
*_device functions take raw device pointers (e.g. torch tensors' data_ptr())."""

from __future__ import annotations

import ctypes as C
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libalaya_b200.so")

u8p, u32p, u64p, f32p = C.POINTER(C.c_uint8), C.POINTER(C.c_uint32), C.POINTER(C.c_uint64), C.POINTER(C.c_float)
METRICS = {"l2": 0, "euclidean": 0, "ip": 1, "cos": 2, "cosine": 2}
QUANTS = {"none": 0, None: 0, "sq8": 1, "sq4": 2}
INDEX_TYPES = {"flat": 0, "hnsw": 1, "nsg": 2, "fusion": 3}

# every symbol include/alaya_b200.h declares (tests check that the library exports all of them)
SYMBOLS = [
    "ab_index_create", "ab_index_destroy", "ab_index_fit", "ab_index_search", "ab_index_batch_search",
    "ab_index_batch_search_with_distance", "ab_index_insert", "ab_index_remove", "ab_index_get_data_by_id",
    "ab_index_get_dim", "ab_index_get_size", "ab_index_save", "ab_index_load", "ab_index_batch_search_device",
    "ab_quant_minmax", "ab_index_fit_ranged", "ab_group_create", "ab_group_destroy", "ab_group_fit", "ab_group_batch_search",
    "ab_group_index", "ab_comm_create", "ab_comm_connect", "ab_comm_batch_search_device", "ab_comm_errors", "ab_comm_destroy",
    "ab_index_fit_spaces", "ab_index_set_graph", "ab_index_build_graph", "ab_index_build_graph_nsg", "ab_index_knng_nndescent", "ab_index_graph_info", "ab_index_get_graph",
    "ab_index_get_codes", "ab_index_search_stats", "ab_index_search_stats_ex", "ab_index_batch_traverse", "ab_distance_pairs", "ab_quant_fit_encode",
    "ab_normalize_rows", "ab_bruteforce_topk", "ab_bruteforce_topk_device", "ab_merge_topk_device", "ab_last_error",
    "ab_device_count", "ab_set_device", "ab_kernel_launch_count", "ab_version", "ab_debug_search_plan",
]


class Params(C.Structure):
    _fields_ = [("index_type", C.c_int32), ("quantization_type", C.c_int32), ("metric", C.c_int32), ("capacity", C.c_uint32),
                ("max_nbrs", C.c_uint32), ("device", C.c_int32)]


_lib = None


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise ImportError(f"{LIB_PATH} is missing: run `python -m alayalite_b200.build`; there is no CPU fallback")
        _lib = C.CDLL(LIB_PATH)
        _lib.ab_last_error.restype = C.c_char_p
        _lib.ab_version.restype = C.c_char_p
        _lib.ab_kernel_launch_count.restype = C.c_uint64
        _lib.ab_index_get_dim.restype = C.c_uint32
        _lib.ab_index_get_size.restype = C.c_uint64
        _lib.ab_index_get_dim.argtypes = [C.c_void_p]
        _lib.ab_index_get_size.argtypes = [C.c_void_p]
        _lib.ab_index_destroy.argtypes = [C.c_void_p]
        _lib.ab_group_destroy.argtypes = [C.c_void_p]
        _lib.ab_comm_destroy.argtypes = [C.c_void_p]
    return _lib


class AbError(RuntimeError):
    def __init__(self, code, msg):
        super().__init__(msg)
        self.code = code


def check(rc):
    if rc != 0:
        raise AbError(rc, lib().ab_last_error().decode())


def _f32(a):
    return np.ascontiguousarray(a, dtype=np.float32)


def _p(a, t):
    return a.ctypes.data_as(t)


def _vp(x):
    return C.c_void_p(int(x)) if x is not None else C.c_void_p(0)


def debug_search_plan(quant, row_chunks, ef, nq, n_rows, sm_count=148, smem_limit=232448):
    """Host-side launch plan of the traversal (no device needed): dict of the six fields of ab_debug_search_plan."""
    out = (C.c_uint64 * 6)()
    check(lib().ab_debug_search_plan(C.c_int32(QUANTS[quant] if (quant is None or isinstance(quant, str)) else quant), C.c_uint32(row_chunks), C.c_uint32(ef),
                                     C.c_uint64(nq), C.c_uint64(n_rows), C.c_int32(sm_count), C.c_uint64(smem_limit), out))
    keys = ("warps_per_cta", "ctas_per_sm", "ctas", "hash_entries", "entry_bytes", "smem_per_cta")
    return dict(zip(keys, [int(x) for x in out]))


def kernel_launch_count() -> int:
    return int(lib().ab_kernel_launch_count())


def device_count() -> int:
    return int(lib().ab_device_count())


class NativeIndex:
    """Thin owner of one ab_index handle."""

    def __init__(self, metric="l2", quant="none", capacity=100000, index_type="hnsw", device=-1, handle=None):
        self.metric, self.quant, self.capacity = metric, quant, capacity
        self._own = handle is None
        if handle is None:
            p = Params(INDEX_TYPES[index_type], QUANTS[quant], METRICS[metric], capacity, 32, device)
            h = C.c_void_p()
            check(lib().ab_index_create(C.byref(p), C.byref(h)))
            self.h = h
        else:
            self.h = C.c_void_p(handle)

    def __del__(self):
        if getattr(self, "_own", False) and getattr(self, "h", None):
            try:
                lib().ab_index_destroy(self.h)
            except TypeError:  # interpreter shutdown: module globals are already gone
                pass
            self.h = None

    @property
    def dim(self):
        return int(lib().ab_index_get_dim(self.h))

    @property
    def size(self):
        return int(lib().ab_index_get_size(self.h))

    def row_bytes(self):
        return self.dim if self.quant == "sq8" else (self.dim + 1) // 2

    def fit(self, data, ef_construction=100, num_threads=1):
        d = _f32(data)
        check(lib().ab_index_fit(self.h, _p(d, f32p), C.c_uint64(d.shape[0]), C.c_uint32(d.shape[1]), C.c_uint32(ef_construction), C.c_uint32(num_threads)))

    def fit_ptr(self, ptr, n, dim, ef_construction=100):
        """Rows behind a raw pointer: host memory, or device memory on the index's device (staged in place)."""
        check(lib().ab_index_fit(self.h, C.cast(_vp(ptr), f32p), C.c_uint64(n), C.c_uint32(dim), C.c_uint32(ef_construction), C.c_uint32(1)))

    def fit_spaces(self, data):
        d = _f32(data)
        check(lib().ab_index_fit_spaces(self.h, _p(d, f32p), C.c_uint64(d.shape[0]), C.c_uint32(d.shape[1])))

    def build_graph(self, ef_construction=100):
        check(lib().ab_index_build_graph(self.h, C.c_uint32(ef_construction)))

    def build_graph_nsg(self, ef_construction=100, nndescent_iters=10):
        check(lib().ab_index_build_graph_nsg(self.h, C.c_uint32(ef_construction), C.c_uint32(nndescent_iters)))

    def knng_nndescent(self, k=32, iters=10):
        out = np.zeros((self.size, 32), np.uint32)
        check(lib().ab_index_knng_nndescent(self.h, C.c_uint32(k), C.c_uint32(iters), _p(out, u32p)))
        return out

    def set_graph(self, adj, levels, lists_flat, ep):
        adj = np.ascontiguousarray(adj, np.uint32)
        lists_flat = np.ascontiguousarray(lists_flat, np.uint32)
        lv = None if levels is None else np.ascontiguousarray(levels, np.uint32)
        check(lib().ab_index_set_graph(self.h, _p(adj, u32p), _p(lv, u32p) if lv is not None else None,
                                       _p(lists_flat, u32p) if lists_flat.size else None, C.c_uint64(lists_flat.size), C.c_uint32(ep)))

    def graph(self):
        info = (C.c_uint64 * 5)()
        check(lib().ab_index_graph_info(self.h, info))
        n, R, ep, ll = int(info[0]), int(info[1]), int(info[2]), int(info[3])
        adj = np.zeros((n, R), np.uint32)
        levels = np.zeros(n, np.uint32)
        lists = np.zeros(max(ll, 1), np.uint32)
        check(lib().ab_index_get_graph(self.h, _p(adj, u32p), _p(levels, u32p), _p(lists, u32p)))
        return dict(adj=adj, levels=levels, lists_flat=lists[:ll], ep=ep, R=R, capacity=int(info[4]))

    def codes(self):
        n, d = self.size, self.dim
        codes = np.zeros((n, self.row_bytes()), np.uint8)
        mn, mx = np.zeros(d, np.float32), np.zeros(d, np.float32)
        check(lib().ab_index_get_codes(self.h, _p(codes, u8p), _p(mn, f32p), _p(mx, f32p)))
        return mn, mx, codes

    def get_data_by_id(self, i):
        out = np.zeros(self.dim, np.float32)
        check(lib().ab_index_get_data_by_id(self.h, C.c_uint32(i), _p(out, f32p)))
        return out

    def search(self, query, topk, ef=100):
        q = _f32(query)
        ids = np.zeros(topk, np.uint32)
        check(lib().ab_index_search(self.h, _p(q, f32p), C.c_uint32(topk), C.c_uint32(ef), _p(ids, u32p)))
        return ids

    def batch_search(self, queries, topk, ef=100, num_threads=1, out=None):
        q = _f32(queries)
        ids = out if out is not None else np.zeros((q.shape[0], topk), np.uint32)
        check(lib().ab_index_batch_search(self.h, _p(q, f32p), C.c_uint64(q.shape[0]), C.c_uint32(topk), C.c_uint32(ef), C.c_uint32(num_threads), _p(ids, u32p)))
        return ids

    def batch_search_ptr(self, q_ptr, nq, topk, ef, ids_ptr):
        """Host pointers (e.g. pinned torch tensors); no numpy on the path."""
        check(lib().ab_index_batch_search(self.h, C.cast(_vp(q_ptr), f32p), C.c_uint64(nq), C.c_uint32(topk), C.c_uint32(ef), C.c_uint32(1), C.cast(_vp(ids_ptr), u32p)))

    def batch_search_with_distance(self, queries, topk, ef=100, num_threads=1):
        q = _f32(queries)
        ids = np.zeros((q.shape[0], topk), np.uint32)
        dist = np.zeros((q.shape[0], topk), np.float32)
        has = C.c_int32(0)
        check(lib().ab_index_batch_search_with_distance(self.h, _p(q, f32p), C.c_uint64(q.shape[0]), C.c_uint32(topk), C.c_uint32(ef), C.c_uint32(num_threads),
                                                        _p(ids, u32p), _p(dist, f32p), C.byref(has)))
        return ids, (dist if has.value else np.zeros((0, topk), np.float32))

    def batch_search_device(self, d_queries, nq, topk, ef, d_out_ids, d_out_dist=None, rerank=True, stream=None):
        check(lib().ab_index_batch_search_device(self.h, _vp(d_queries), C.c_uint64(nq), C.c_uint32(topk), C.c_uint32(ef), C.c_int32(int(rerank)),
                                                 _vp(d_out_ids), _vp(d_out_dist), _vp(stream)))

    def traverse(self, queries, ef):
        q = _f32(queries)
        nq = q.shape[0]
        ids = np.zeros((nq, ef), np.uint32)
        dist = np.zeros((nq, ef), np.float32)
        sizes = np.zeros(nq, np.uint32)
        check(lib().ab_index_batch_traverse(self.h, _p(q, f32p), C.c_uint64(nq), C.c_uint32(ef), _p(ids, u32p), _p(dist, f32p), _p(sizes, u32p)))
        return ids, dist, sizes

    def stats(self):
        st = (C.c_uint64 * 8)()
        check(lib().ab_index_search_stats_ex(self.h, st, C.c_uint32(8)))
        return dict(ndc=int(st[0]), hops=int(st[1]), overlay=int(st[2]), spilled=int(st[3]), queries=int(st[4]),
                    bitmap=int(st[5]), overflow_ids=int(st[6]), visited_workspace_bytes=int(st[7]))

    def insert(self, vec, ef=100):
        v = _f32(vec)
        out = C.c_uint32(0)
        check(lib().ab_index_insert(self.h, _p(v, f32p), C.c_uint32(ef), C.byref(out)))
        return -1 if out.value == 0xFFFFFFFF else int(out.value)

    def remove(self, vid):
        check(lib().ab_index_remove(self.h, C.c_uint32(vid)))

    def save(self, index_path, data_path, quant_path=""):
        check(lib().ab_index_save(self.h, index_path.encode(), data_path.encode(), quant_path.encode()))

    def load(self, index_path, data_path, quant_path=""):
        check(lib().ab_index_load(self.h, index_path.encode(), data_path.encode(), quant_path.encode()))


def distance_pairs(metric, quant, x, y, mn=None, mx=None, dim=None):
    if QUANTS[quant] == 0:
        x, y = _f32(x), _f32(y)
        dim = x.shape[1]
    else:
        x, y = np.ascontiguousarray(x, np.uint8), np.ascontiguousarray(y, np.uint8)
        mn, mx = _f32(mn), _f32(mx)
        dim = mn.shape[0] if dim is None else dim
    out = np.zeros(x.shape[0], np.float32)
    check(lib().ab_distance_pairs(C.c_int32(METRICS[metric]), C.c_int32(QUANTS[quant]), _vp(x.ctypes.data), _vp(y.ctypes.data), C.c_uint64(x.shape[0]),
                                  C.c_uint32(dim), _p(mn, f32p) if mn is not None else None, _p(mx, f32p) if mx is not None else None, _p(out, f32p)))
    return out


def quant_fit_encode(quant, data):
    d = _f32(data)
    n, dim = d.shape
    rb = dim if quant == "sq8" else (dim + 1) // 2
    mn, mx = np.zeros(dim, np.float32), np.zeros(dim, np.float32)
    codes = np.zeros((n, rb), np.uint8)
    check(lib().ab_quant_fit_encode(C.c_int32(QUANTS[quant]), _p(d, f32p), C.c_uint64(n), C.c_uint32(dim), _p(mn, f32p), _p(mx, f32p), _p(codes, u8p)))
    return mn, mx, codes


def normalize_rows(data):
    d = _f32(data).copy()
    check(lib().ab_normalize_rows(_p(d, f32p), C.c_uint64(d.shape[0]), C.c_uint32(d.shape[1])))
    return d


def bruteforce_topk(metric, data, queries, k):
    d, q = _f32(data), _f32(queries)
    ids = np.zeros((q.shape[0], k), np.uint32)
    dist = np.zeros((q.shape[0], k), np.float32)
    check(lib().ab_bruteforce_topk(C.c_int32(METRICS[metric]), _p(d, f32p), C.c_uint64(d.shape[0]), C.c_uint32(d.shape[1]), _p(q, f32p),
                                   C.c_uint64(q.shape[0]), C.c_uint32(k), _p(ids, u32p), _p(dist, f32p)))
    return ids, dist


def bruteforce_topk_device(metric, d_data, n, dim, d_queries, nq, k, d_out_ids, d_out_dist=None, stream=None):
    check(lib().ab_bruteforce_topk_device(C.c_int32(METRICS[metric]), _vp(d_data), C.c_uint64(n), C.c_uint32(dim), _vp(d_queries), C.c_uint64(nq),
                                          C.c_uint32(k), _vp(d_out_ids), _vp(d_out_dist), _vp(stream)))


def merge_topk_device(d_dist, d_ids, lists, nq, k, d_out_ids, d_out_dist=None, stream=None):
    check(lib().ab_merge_topk_device(_vp(d_dist), _vp(d_ids), C.c_uint32(lists), C.c_uint64(nq), C.c_uint32(k), _vp(d_out_ids), _vp(d_out_dist), _vp(stream)))


# ---- multi-GPU (include/alaya_b200.h "Multi-GPU") -------------------------------------------------------
COMM_HANDLE_BYTES = 128


def quant_minmax(metric, data):
    """Per-dimension (min, max) of `data` the way the SQ quantiser sees the rows (COS: normalised first)."""
    d = _f32(data)
    mn, mx = np.zeros(d.shape[1], np.float32), np.zeros(d.shape[1], np.float32)
    check(lib().ab_quant_minmax(C.c_int32(METRICS[metric]), _p(d, f32p), C.c_uint64(d.shape[0]), C.c_uint32(d.shape[1]), _p(mn, f32p), _p(mx, f32p)))
    return mn, mx


def quant_minmax_ptr(metric, ptr, n, dim):
    mn, mx = np.zeros(dim, np.float32), np.zeros(dim, np.float32)
    check(lib().ab_quant_minmax(C.c_int32(METRICS[metric]), C.cast(_vp(ptr), f32p), C.c_uint64(n), C.c_uint32(dim), _p(mn, f32p), _p(mx, f32p)))
    return mn, mx


def fit_ranged(index: NativeIndex, data_or_ptr, n=None, dim=None, ef_construction=100, mn=None, mx=None):
    """ab_index_fit with a given quantiser range (shards of one logical index share the global range)."""
    if n is None:
        d = _f32(data_or_ptr)
        ptr, n, dim = _p(d, f32p), d.shape[0], d.shape[1]
    else:
        ptr = C.cast(_vp(data_or_ptr), f32p)
    pmn = _p(_f32(mn), f32p) if mn is not None else None
    pmx = _p(_f32(mx), f32p) if mx is not None else None
    check(lib().ab_index_fit_ranged(index.h, ptr, C.c_uint64(n), C.c_uint32(dim), C.c_uint32(ef_construction), pmn, pmx))


class Group:
    """ab_group: a shards x replicas grid of GPUs driven by this process."""

    def __init__(self, metric="l2", quant="none", capacity=100000, devices=(0,), n_shards=1, n_replicas=1, index_type="hnsw"):
        self.metric, self.quant, self.n_shards, self.n_replicas = metric, quant, n_shards, n_replicas
        assert len(devices) == n_shards * n_replicas
        p = Params(INDEX_TYPES[index_type], QUANTS[quant], METRICS[metric], capacity, 32, -1)
        devs = (C.c_int32 * len(devices))(*[int(d) for d in devices])
        h = C.c_void_p()
        check(lib().ab_group_create(C.byref(p), devs, C.c_uint32(n_shards), C.c_uint32(n_replicas), C.byref(h)))
        self.h = h

    def __del__(self):
        if getattr(self, "h", None):
            try:
                lib().ab_group_destroy(self.h)
            except TypeError:
                pass
            self.h = None

    def fit(self, data, ef_construction=100):
        d = _f32(data)
        check(lib().ab_group_fit(self.h, _p(d, f32p), C.c_uint64(d.shape[0]), C.c_uint32(d.shape[1]), C.c_uint32(ef_construction)))

    def batch_search(self, queries, topk, ef=100, with_distance=False):
        q = _f32(queries)
        ids = np.zeros((q.shape[0], topk), np.uint32)
        dist = np.zeros((q.shape[0], topk), np.float32) if with_distance else None
        check(lib().ab_group_batch_search(self.h, _p(q, f32p), C.c_uint64(q.shape[0]), C.c_uint32(topk), C.c_uint32(ef), _p(ids, u32p),
                                          _p(dist, f32p) if with_distance else None))
        return (ids, dist) if with_distance else ids

    def batch_search_ptr(self, q_ptr, nq, topk, ef, ids_ptr, dist_ptr=None):
        check(lib().ab_group_batch_search(self.h, C.cast(_vp(q_ptr), f32p), C.c_uint64(nq), C.c_uint32(topk), C.c_uint32(ef),
                                          C.cast(_vp(ids_ptr), u32p), C.cast(_vp(dist_ptr), f32p) if dist_ptr else None))

    def shard(self, shard=0, replica=0):
        """The ab_index of one shard (owned by the group)."""
        h = C.c_void_p()
        check(lib().ab_group_index(self.h, C.c_uint32(shard), C.c_uint32(replica), C.byref(h)))
        return NativeIndex(self.metric, self.quant, 0, handle=h.value)


class Comm:
    """ab_comm: this process's shard of a replica group (one process per GPU)."""

    def __init__(self, index: NativeIndex, rank, world, id_offset, max_nq, max_topk):
        self.index, self.rank, self.world = index, rank, world
        buf = (C.c_uint8 * COMM_HANDLE_BYTES)()
        h = C.c_void_p()
        check(lib().ab_comm_create(index.h, C.c_uint32(rank), C.c_uint32(world), C.c_uint32(id_offset), C.c_uint64(max_nq), C.c_uint32(max_topk),
                                   C.byref(h), buf))
        self.h, self.handle = h, bytes(buf)

    def connect(self, handles):
        blob = b"".join(handles)
        assert len(blob) == self.world * COMM_HANDLE_BYTES
        arr = (C.c_uint8 * len(blob)).from_buffer_copy(blob)
        check(lib().ab_comm_connect(self.h, arr))

    def owned(self, nq):
        return (nq - self.rank + self.world - 1) // self.world if nq > self.rank else 0

    def batch_search_device(self, d_queries, nq, topk, ef, d_out_ids, d_out_dist=None, stream=None):
        check(lib().ab_comm_batch_search_device(self.h, _vp(d_queries), C.c_uint64(nq), C.c_uint32(topk), C.c_uint32(ef), _vp(d_out_ids),
                                                _vp(d_out_dist), _vp(stream)))

    def errors(self):
        out = C.c_uint32(0)
        check(lib().ab_comm_errors(self.h, C.byref(out)))
        return int(out.value)

    def close(self):
        if getattr(self, "h", None):
            lib().ab_comm_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except TypeError:
            pass
