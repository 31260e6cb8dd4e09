"""TEST INFRASTRUCTURE ONLY: ctypes access to the two CPU checkers.

* ``Checker("orc")``  -> oracle/liboracle.so, our restatement of the reference algorithm
  (oracle/restate.cpp).
* ``Checker("ref")``  -> oracle/_ref/libref_oracle.so, the unmodified reference headers compiled in
  place (oracle/ref_harness.cpp); present here, and on the GPU box as a prebuilt file.

Both expose the same function set (prefix ``orc_`` / ``ref_``), so the same test body can run
against either. Nothing under alayalite_b200/ may import this module: only tests/, smoke() and
bench.py's cpu_baseline / --impl reference legs do.
"""

from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ORC_SO = os.path.join(HERE, "liboracle.so")
REF_SO = os.path.join(HERE, "_ref", "libref_oracle.so")

u8p = C.POINTER(C.c_uint8)
u32p = C.POINTER(C.c_uint32)
u64p = C.POINTER(C.c_uint64)
f32p = C.POINTER(C.c_float)
i32p = C.POINTER(C.c_int)

METRICS = {"l2": 0, "euclidean": 0, "ip": 1, "cos": 2, "cosine": 2}
QUANTS = {"none": 0, None: 0, "sq8": 1, "sq4": 2}


def build(ref: bool = True) -> None:
    """Compile the checkers (building the checker is not using it)."""
    targets = ["liboracle.so"] + (["ref"] if ref else [])
    subprocess.run(["make", "-C", HERE] + targets, check=True, stdout=subprocess.DEVNULL)


def have_ref() -> bool:
    return os.path.exists(REF_SO)


def _p(a, t):
    return a.ctypes.data_as(t)


def _f32(a):
    return np.ascontiguousarray(a, dtype=np.float32)


class Checker:
    """One of the two CPU implementations behind a uniform numpy interface."""

    def __init__(self, which: str = "orc"):
        assert which in ("orc", "ref")
        self.which = which
        path = ORC_SO if which == "orc" else REF_SO
        if not os.path.exists(path):
            raise FileNotFoundError(f"{path} missing; run `make -C oracle`")
        self.lib = C.CDLL(path)
        self._sig()

    def _fn(self, name):
        return getattr(self.lib, f"{self.which}_{name}")

    def _sig(self):
        L = self._fn
        L("index_create").restype = C.c_void_p
        L("index_create").argtypes = [C.c_int, C.c_int, C.c_uint32, C.c_uint32]
        L("index_destroy").argtypes = [C.c_void_p]
        for n in ("l2_sqr", "ip_sqr"):
            L(n).restype = C.c_float
            L(n).argtypes = [f32p, f32p, C.c_size_t]
        for n in ("l2_sqr_sq8", "ip_sqr_sq8", "l2_sqr_sq4", "ip_sqr_sq4"):
            L(n).restype = C.c_float
            L(n).argtypes = [u8p, u8p, C.c_size_t, f32p, f32p]
        for n in ("sq8_quantize", "sq4_quantize"):
            L(n).restype = C.c_uint8
            L(n).argtypes = [C.c_float, C.c_float, C.c_float]
        L("index_space_distance").restype = C.c_float
        L("index_space_distance").argtypes = [C.c_void_p, C.c_uint32, C.c_uint32]
        L("index_insert").restype = C.c_int64
        L("index_insert").argtypes = [C.c_void_p, f32p, C.c_uint32]
        L("normalize").argtypes = [f32p, C.c_size_t]
        if self.which == "ref":
            L("simd_level").restype = C.c_char_p
            L("last_error").restype = C.c_char_p
            for n in ("l2_sqr_generic", "ip_sqr_generic"):
                L(n).restype = C.c_float
                L(n).argtypes = [f32p, f32p, C.c_size_t]
            for n in ("l2_sqr_sq8_generic", "ip_sqr_sq8_generic", "l2_sqr_sq4_generic", "ip_sqr_sq4_generic"):
                L(n).restype = C.c_float
                L(n).argtypes = [u8p, u8p, C.c_size_t, f32p, f32p]
        else:
            L("calc_recall").restype = C.c_float

    # ---- kernels -------------------------------------------------------------------------------
    def dist(self, name, x, y, mn=None, mx=None, generic=False):
        fn = self._fn(name + ("_generic" if generic else ""))
        if mn is None:
            x, y = _f32(x), _f32(y)
            return float(fn(_p(x, f32p), _p(y, f32p), x.shape[0]))
        x = np.ascontiguousarray(x, dtype=np.uint8)
        y = np.ascontiguousarray(y, dtype=np.uint8)
        mn, mx = _f32(mn), _f32(mx)
        return float(fn(_p(x, u8p), _p(y, u8p), mn.shape[0], _p(mn, f32p), _p(mx, f32p)))

    def quantize(self, bits, v, mn, mx):
        return int(self._fn(f"sq{bits}_quantize")(float(v), float(mn), float(mx)))

    def fit_encode(self, bits, data):
        data = _f32(data)
        n, d = data.shape
        rb = d if bits == 8 else (d + 1) // 2
        mn = np.empty(d, np.float32)
        mx = np.empty(d, np.float32)
        codes = np.zeros((n, rb), np.uint8)
        self._fn(f"sq{bits}_fit_encode")(_p(data, f32p), C.c_size_t(n), C.c_uint32(d), _p(mn, f32p), _p(mx, f32p), _p(codes, u8p))
        return mn, mx, codes

    def normalize(self, v):
        v = _f32(v).copy()
        self._fn("normalize")(_p(v, f32p), v.shape[0])
        return v

    def find_exact_gt(self, queries, data, topk, threads=1):
        queries, data = _f32(queries), _f32(data)
        ids = np.zeros((queries.shape[0], topk), np.uint32)
        self._fn("find_exact_gt")(_p(queries, f32p), C.c_uint32(queries.shape[0]), _p(data, f32p), C.c_uint32(data.shape[0]),
                                  C.c_uint32(data.shape[1]), C.c_uint32(topk), C.c_uint32(threads), _p(ids, u32p))
        return ids

    def bruteforce_topk(self, metric, queries, data, topk, threads=1):
        assert self.which == "orc"
        queries, data = _f32(queries), _f32(data)
        ids = np.zeros((queries.shape[0], topk), np.uint32)
        dist = np.zeros((queries.shape[0], topk), np.float32)
        self.lib.orc_bruteforce_topk(C.c_int(METRICS[metric]), _p(queries, f32p), C.c_uint32(queries.shape[0]), _p(data, f32p),
                                     C.c_uint32(data.shape[0]), C.c_uint32(data.shape[1]), C.c_uint32(topk), C.c_uint32(threads),
                                     _p(ids, u32p), _p(dist, f32p))
        return ids, dist

    def linear_pool_replay(self, n, capacity, ops):
        """ops: list of ("insert", id, dist) / ("pop",). Returns (popped ids, final ids, final dists)."""
        op_ids = np.array([0xFFFFFFFF if o[0] == "pop" else o[1] for o in ops], np.uint32)
        op_d = np.array([0.0 if o[0] == "pop" else o[2] for o in ops], np.float32)
        popped = np.zeros(len(ops) + 1, np.uint32)
        fin_ids = np.zeros(capacity + 1, np.uint32)
        fin_d = np.zeros(capacity + 1, np.float32)
        npop = C.c_uint32(0)
        fsz = C.c_uint32(0)
        self._fn("linear_pool_replay")(C.c_uint32(n), C.c_int(capacity), _p(op_ids, u32p), _p(op_d, f32p), C.c_uint32(len(ops)),
                                       _p(popped, u32p), C.byref(npop), _p(fin_ids, u32p), _p(fin_d, f32p), C.byref(fsz))
        return popped[: npop.value].copy(), fin_ids[: fsz.value].copy(), fin_d[: fsz.value].copy()

    def hnsw_levels(self, count, max_edge=16):
        assert self.which == "orc"
        out = np.zeros(count, np.uint32)
        self.lib.orc_hnsw_levels(C.c_uint32(count), C.c_uint32(max_edge), _p(out, u32p))
        return out

    def simd_level(self):
        return self.lib.ref_simd_level().decode() if self.which == "ref" else "scalar"

    def index(self, metric="l2", quant="none", capacity=100000, dim=0):
        return CheckerIndex(self, metric, quant, capacity, dim)


class CheckerIndex:
    """Mirror of PyIndexInterface (python/src/pybind.cpp:105-147) on a CPU checker."""

    def __init__(self, chk: Checker, metric, quant, capacity, dim):
        self.chk, self.dim, self.capacity = chk, dim, capacity
        self.metric, self.quant = metric, quant
        self.h = C.c_void_p(chk._fn("index_create")(METRICS[metric], QUANTS[quant], capacity, dim))
        self.n = 0

    def __del__(self):
        try:
            self.chk._fn("index_destroy")(self.h)
        except Exception:  # pylint: disable=broad-except
            pass

    def _check(self, rc):
        if rc != 0:
            msg = self.chk.lib.ref_last_error().decode() if self.chk.which == "ref" else "oracle error"
            raise RuntimeError(msg)

    def fit(self, data, ef_construction=100, num_threads=1):
        """NOTE the reference normalises `data` in place for COS (raw_space.hpp:131-139); pass a copy."""
        data = _f32(data)
        self.n = data.shape[0]
        self._data = data
        if self.chk.which == "ref":
            self._check(self.chk.lib.ref_index_fit(self.h, _p(data, f32p), C.c_uint32(self.n), C.c_uint32(ef_construction), C.c_uint32(num_threads)))
            return None
        calls = C.c_uint64(0)
        self._check(self.chk.lib.orc_index_fit(self.h, _p(data, f32p), C.c_uint32(self.n), C.c_uint32(ef_construction), C.byref(calls)))
        return calls.value

    def fit_nsg(self, data, ef_construction=100):
        """The reference's NSGBuilder (NN-Descent + link + tree_grow); reference only."""
        assert self.chk.which == "ref"
        data = _f32(data)
        self.n = data.shape[0]
        self._data = data
        self._check(self.chk.lib.ref_index_fit_nsg(self.h, _p(data, f32p), C.c_uint32(self.n), C.c_uint32(ef_construction)))

    def knng(self, data, k=32):
        """The reference's NndescentImpl k-NN graph: ids [n, k]; reference only."""
        assert self.chk.which == "ref"
        data = _f32(data)
        out = np.zeros((data.shape[0], k), np.uint32)
        self._check(self.chk.lib.ref_index_knng(self.h, _p(data, f32p), C.c_uint32(data.shape[0]), C.c_uint32(k), _p(out, u32p)))
        return out

    def fit_spaces(self, data):
        assert self.chk.which == "orc"
        data = _f32(data)
        self.n = data.shape[0]
        self._check(self.chk.lib.orc_index_fit_spaces(self.h, _p(data, f32p), C.c_uint32(self.n)))

    def set_graph(self, adj, levels, lists_flat, ep, has_overlay=True):
        assert self.chk.which == "orc"
        adj = np.ascontiguousarray(adj, np.uint32)
        levels = np.ascontiguousarray(levels, np.uint32)
        lists_flat = np.ascontiguousarray(lists_flat, np.uint32)
        if lists_flat.size == 0:
            lists_flat = np.zeros(1, np.uint32)
        self._check(self.chk.lib.orc_index_set_graph(self.h, _p(adj, u32p), _p(levels, u32p), _p(lists_flat, u32p), C.c_uint32(ep), C.c_int(int(has_overlay))))

    def graph(self):
        """-> dict(adj [n,R], levels [n], lists_flat, ep, max_nodes, has_overlay)."""
        info = np.zeros(5, np.uint32)
        self._check(self.chk._fn("index_graph_info")(self.h, _p(info, u32p)))
        n, R = int(info[4]), int(info[1])
        self.n = n
        adj = np.zeros((n, R), np.uint32)
        self.chk._fn("index_graph_rows")(self.h, C.c_uint32(0), C.c_uint32(n), _p(adj, u32p))
        levels = np.zeros(n, np.uint32)
        lists = []
        if info[3]:
            self.chk._fn("index_overlay_levels")(self.h, C.c_uint32(n), _p(levels, u32p))
            for i in np.nonzero(levels)[0]:
                buf = np.zeros(int(levels[i]) * R, np.uint32)
                self.chk._fn("index_overlay_list")(self.h, C.c_uint32(int(i)), _p(buf, u32p))
                lists.append(buf)
        flat = np.concatenate(lists) if lists else np.zeros(0, np.uint32)
        return dict(adj=adj, levels=levels, lists_flat=flat, ep=int(info[2]), max_nodes=int(info[0]), has_overlay=bool(info[3]), R=R)

    def raw_rows(self, n=None):
        n = self.n if n is None else n
        out = np.zeros((n, self.dim), np.float32)
        for i in range(n):
            self.chk._fn("index_raw_row")(self.h, C.c_uint32(i), _p(out[i], f32p))
        return out

    def codes(self, n=None):
        n = self.n if n is None else n
        rb = self.dim if self.quant == "sq8" else (self.dim + 1) // 2
        out = np.zeros((n, rb), np.uint8)
        for i in range(n):
            self._check(self.chk._fn("index_code_row")(self.h, C.c_uint32(i), _p(out[i], u8p), C.c_uint32(rb)))
        return out

    def minmax(self):
        mn = np.zeros(self.dim, np.float32)
        mx = np.zeros(self.dim, np.float32)
        self._check(self.chk._fn("index_quant_minmax")(self.h, _p(mn, f32p), _p(mx, f32p)))
        return mn, mx

    def space_distance(self, i, j):
        return float(self.chk._fn("index_space_distance")(self.h, C.c_uint32(i), C.c_uint32(j)))

    def search(self, query, topk, ef=100):
        q = _f32(query).copy()
        ids = np.zeros(topk, np.uint32)
        self._check(self.chk._fn("index_search")(self.h, _p(q, f32p), C.c_uint32(topk), C.c_uint32(ef), _p(ids, u32p)))
        return ids

    def search_updated(self, query, topk, ef=100):
        """GraphSearchJob::search_solo_updated (reference only): pool order, no re-rank."""
        assert self.chk.which == "ref"
        q = _f32(query).copy()
        ids = np.zeros(topk, np.uint32)
        self._check(self.chk.lib.ref_index_search_updated(self.h, _p(q, f32p), C.c_uint32(topk), C.c_uint32(ef), _p(ids, u32p)))
        return ids

    def batch_search(self, queries, topk, ef=100, num_threads=1, solo=False):
        """solo=True (ref only): std::thread x search_solo instead of the coroutine Scheduler."""
        q = _f32(queries).copy()  # the reference normalises COS queries in place
        ids = np.zeros((q.shape[0], topk), np.uint32)
        name = "index_batch_search_solo" if (solo and self.chk.which == "ref") else "index_batch_search"
        self._check(self.chk._fn(name)(self.h, _p(q, f32p), C.c_uint32(q.shape[0]), C.c_uint32(topk), C.c_uint32(ef), C.c_uint32(num_threads), _p(ids, u32p)))
        return ids

    def batch_search_inplace(self, q, topk, ef, num_threads, solo=False):
        """Timing helper: no copies, caller owns a contiguous float32 array."""
        ids = np.zeros((q.shape[0], topk), np.uint32)
        name = "index_batch_search_solo" if (solo and self.chk.which == "ref") else "index_batch_search"
        self._check(self.chk._fn(name)(self.h, _p(q, f32p), C.c_uint32(q.shape[0]), C.c_uint32(topk), C.c_uint32(ef), C.c_uint32(num_threads), _p(ids, u32p)))
        return ids

    def batch_search_with_distance(self, queries, topk, ef=100, num_threads=1):
        q = _f32(queries).copy()
        ids = np.zeros((q.shape[0], topk), np.uint32)
        dist = np.zeros((q.shape[0], topk), np.float32)
        has = C.c_int(0)
        self._check(self.chk._fn("index_batch_search_with_distance")(self.h, _p(q, f32p), C.c_uint32(q.shape[0]), C.c_uint32(topk), C.c_uint32(ef),
                                                                     C.c_uint32(num_threads), _p(ids, u32p), _p(dist, f32p), C.byref(has)))
        return ids, (dist if has.value else np.zeros((0, topk), np.float32))

    def traverse(self, query, ef):
        """Search-space traversal of one query: (pool ids, pool dists, stats{ndc,hops,overlay})."""
        q = _f32(query).copy()
        ids = np.zeros(ef + 1, np.uint32)
        dist = np.zeros(ef + 1, np.float32)
        size = C.c_uint32(0)
        stats = np.zeros(3, np.uint64)
        self._check(self.chk._fn("index_traverse")(self.h, _p(q, f32p), C.c_uint32(ef), _p(ids, u32p), _p(dist, f32p), C.byref(size), _p(stats, u64p)))
        return ids[: size.value].copy(), dist[: size.value].copy(), dict(ndc=int(stats[0]), hops=int(stats[1]), overlay=int(stats[2]))

    def insert(self, vec, ef=100):
        v = _f32(vec).copy()
        r = int(self.chk._fn("index_insert")(self.h, _p(v, f32p), C.c_uint32(ef)))
        if self.chk.which == "ref" and r == 0xFFFFFFFF:
            r = -1
        return r

    def remove(self, vid):
        self._check(self.chk._fn("index_remove")(self.h, C.c_uint32(vid)))

    def save(self, index_path, data_path, quant_path=""):
        assert self.chk.which == "ref"
        self._check(self.chk.lib.ref_index_save(self.h, index_path.encode(), data_path.encode(), quant_path.encode()))

    def load(self, index_path, data_path, quant_path=""):
        assert self.chk.which == "ref"
        self._check(self.chk.lib.ref_index_load(self.h, index_path.encode(), data_path.encode(), quant_path.encode()))
        info = np.zeros(5, np.uint32)
        self.chk.lib.ref_index_graph_info(self.h, _p(info, u32p))
        self.n = int(info[4])
