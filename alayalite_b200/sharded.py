"""ShardedIndex: one process per GPU; a grid of id-range shards x query replicas.

The reference has no multi-device mode; this is the one addition. The work is done by libalaya_b200.so
(`ab_comm_*`, include/alaya_b200.h): rank s of a replica group owns the contiguous id range
[offset_s, offset_s + n_s), builds its own graph over it with local ids, and answers every query of its group on
its shard. The kernel that produces a query's final list stores it (global ids) into the mailbox of the query's
OWNER shard -- query i belongs to shard i % S -- over NVLink peer memory, and the last shard to arrive merges the S
lists by (distance, id) on the spot. No collective, no barrier, no merge kernel; what a CPU merge of the reference's
per-shard results would give.

torch.distributed is plumbing only: it moves the 128-byte mailbox handles at set-up, the d floats of the global SQ
range at fit time, and -- for callers who want every row on every rank -- the final all-gather of owned rows. On
hosts without CUDA (the world-size-2 gloo tests of this logic) or with AB_PEER_MEMORY=0 the per-shard lists go
through all_gather_into_tensor and a merge function instead; both routes return the same rows.
"""

from __future__ import annotations

import os

import numpy as np


def shard_range(n: int, shards: int, s: int):
    """Rows of shard s when n rows are split into contiguous ranges (the rule of ab_group_fit)."""
    per = (n + shards - 1) // shards
    return s * per, min(n, (s + 1) * per)


def global_quantizer_range(metric, rows_ptr_or_array, n=None, dim=None, group=None):
    """Per-dimension min / max over ALL shards: each rank's own range (ab_quant_minmax) all-reduced with MIN / MAX, so
    that every shard encodes with the range of the single index (its codes then equal the single index's codes)."""
    import torch
    import torch.distributed as dist

    from . import _cabi

    if n is None:
        mn, mx = _cabi.quant_minmax(metric, rows_ptr_or_array)
    else:
        mn, mx = _cabi.quant_minmax_ptr(metric, rows_ptr_or_array, n, dim)
    if dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1:
        dev = torch.device("cuda", torch.cuda.current_device()) if dist.get_backend(group) == "nccl" else torch.device("cpu")
        tmn, tmx = torch.from_numpy(mn).to(dev), torch.from_numpy(mx).to(dev)
        dist.all_reduce(tmn, op=dist.ReduceOp.MIN, group=group)
        dist.all_reduce(tmx, op=dist.ReduceOp.MAX, group=group)
        mn, mx = tmn.cpu().numpy(), tmx.cpu().numpy()
    return mn, mx


class ShardedIndex:
    def __init__(self, local_index, id_offset: int, world: int, rank: int = None, group=None, merge_fn=None, search_fn=None,
                 peer_memory=True, max_nq: int = 16384, max_topk: int = 128):
        """local_index: _cabi.NativeIndex of this rank's shard (already fitted). world / rank / group describe the
        REPLICA GROUP this rank shards with (group=None: the default process group).
        merge_fn / search_fn are injection points for the CPU (gloo) tests of the host logic."""
        import torch
        import torch.distributed as dist

        self.torch, self.dist = torch, dist
        self.ix, self.offset, self.world, self.group = local_index, int(id_offset), int(world), group
        self.rank = int(rank) if rank is not None else (dist.get_rank(group) if (world > 1 and dist.is_initialized()) else 0)
        self._merge = merge_fn or self._merge_device
        self._search = search_fn or self._search_device
        self._buf = {}
        self._comm = None
        self.max_nq, self.max_topk = int(max_nq), int(max_topk)
        self.peer_memory = (bool(peer_memory) and search_fn is None and merge_fn is None and self.world > 1
                            and os.environ.get("AB_PEER_MEMORY", "1") != "0")
        self.exchange_kind = "none" if self.world == 1 else "nccl all_gather + merge kernel"
        self.peer_error = None

    # ---- owner bookkeeping ------------------------------------------------------------------------------
    def owned(self, nq):
        """Queries whose merged rows this rank produces: rank, rank + world, ..."""
        return np.arange(self.rank, nq, self.world)

    # ---- set-up of the fused exchange -------------------------------------------------------------------
    def _communicator(self):
        if self._comm is None and self.peer_memory:
            from . import _cabi

            try:
                comm = _cabi.Comm(self.ix, self.rank, self.world, self.offset, self.max_nq, self.max_topk)
                handles = [None] * self.world
                self.dist.all_gather_object(handles, comm.handle, group=self.group)
                comm.connect(handles)
                self._comm = comm
                self.exchange_kind = "fused: peer stores from the search epilogue + last-arriver merge (ab_comm)"
            except Exception as e:  # noqa: BLE001  (no peer access here: keep the collective route)
                self.peer_memory = False
                self.peer_error = repr(e)
        return self._comm

    # ---- the two device steps of the collective route ---------------------------------------------------
    def _search_device(self, d_queries, nq, topk, ef, d_ids, d_dist, stream):
        # rerank = 3: SQ re-rank + slots a small shard cannot fill marked (0xFFFFFFFF, FLT_MAX) instead of the reference's
        # zero padding, which would otherwise enter the merge as row 0 of this shard at distance 0
        self.ix.batch_search_device(d_queries.data_ptr(), nq, topk, ef, d_ids.data_ptr(), d_dist.data_ptr(), 3 if self.world > 1 else 1, stream)

    def _merge_device(self, g_dist, g_ids, nq, topk, out_ids, out_dist, stream):
        from . import _cabi

        _cabi.merge_topk_device(g_dist.data_ptr(), g_ids.data_ptr(), self.world, nq, topk, out_ids.data_ptr(), out_dist.data_ptr(), stream)

    def _buffers(self, nq, topk, device):
        key = (nq, topk, str(device))
        if key not in self._buf:
            t = self.torch
            rows = (nq + self.world - 1) // self.world
            self._buf[key] = dict(
                ids=t.empty((nq, topk), dtype=t.int32, device=device), dist=t.empty((nq, topk), dtype=t.float32, device=device),
                g_ids=t.empty((self.world, nq, topk), dtype=t.int32, device=device), g_dist=t.empty((self.world, nq, topk), dtype=t.float32, device=device),
                m_ids=t.empty((nq, topk), dtype=t.int32, device=device), m_dist=t.empty((nq, topk), dtype=t.float32, device=device),
                o_ids=t.empty((rows, topk), dtype=t.int32, device=device), o_dist=t.empty((rows, topk), dtype=t.float32, device=device))
        return self._buf[key]

    def _stream_ctx(self, stream):
        """torch ops of a step must run on the stream the library's kernels were enqueued on."""
        t = self.torch
        if stream is None or not t.cuda.is_available():
            import contextlib

            return contextlib.nullcontext(), None
        if isinstance(stream, t.cuda.Stream):
            return t.cuda.stream(stream), stream.cuda_stream
        return t.cuda.stream(t.cuda.ExternalStream(int(stream))), int(stream)

    def exchange(self, local_ids, local_dist, b, stream=None):
        """Collective route: local (already global-id) lists -> all ranks hold the merged [nq, topk] result."""
        nq, topk = local_ids.shape
        if self.world == 1:
            return local_ids, local_dist
        ctx, raw = self._stream_ctx(stream)
        with ctx:
            # [world, nq, topk] filled rank by rank == concatenation along dim 0 of the [nq, topk] inputs
            self.dist.all_gather_into_tensor(b["g_ids"].view(self.world * nq, topk), local_ids, group=self.group)
            self.dist.all_gather_into_tensor(b["g_dist"].view(self.world * nq, topk), local_dist, group=self.group)
            self._merge(b["g_dist"], b["g_ids"], nq, topk, b["m_ids"], b["m_dist"], raw)
        return b["m_ids"], b["m_dist"]

    # ---- search -------------------------------------------------------------------------------------------
    def batch_search_device(self, d_queries, topk, ef, stream=None, out_ids=None, out_dist=None):
        """d_queries [nq, dim] on this rank's device (the same queries on every rank of the group).
        Returns (ids, dist) of the queries this rank OWNS -- rows rank, rank + world, ... of the merged result, global
        ids -- as [len(owned(nq)), topk] tensors (device tensors, or `out_ids` / `out_dist` if given: device or pinned
        host tensors written by the library). Asynchronous on `stream` (a torch Stream or a raw handle)."""
        t = self.torch
        nq = d_queries.shape[0]
        device = d_queries.device
        ctx, raw = self._stream_ctx(stream)
        b = self._buffers(nq, topk, device)
        n_owned = len(range(self.rank, nq, self.world))
        o_ids = out_ids if out_ids is not None else b["o_ids"]
        o_dist = out_dist if out_dist is not None else b["o_dist"]
        if self.world == 1:
            self._search(d_queries, nq, topk, ef, o_ids, o_dist, raw)
            return o_ids[:n_owned], o_dist[:n_owned]
        comm = self._communicator() if device.type == "cuda" else None
        if comm is not None:
            comm.batch_search_device(d_queries.data_ptr(), nq, topk, ef, o_ids.data_ptr(), o_dist.data_ptr(), raw)
            return o_ids[:n_owned], o_dist[:n_owned]
        with ctx:
            self._search(d_queries, nq, topk, ef, b["ids"], b["dist"], raw)
            b["ids"].add_(self.offset * (b["ids"] >= 0).to(b["ids"].dtype))  # local -> global ids; unfilled slots (-1) stay
            m_ids, m_dist = self.exchange(b["ids"], b["dist"], b, stream)
            o_ids[:n_owned].copy_(m_ids[self.rank::self.world], non_blocking=True)
            o_dist[:n_owned].copy_(m_dist[self.rank::self.world], non_blocking=True)
        return o_ids[:n_owned], o_dist[:n_owned]

    def batch_search(self, queries, topk, ef, device=None, stream=None):
        """Host queries in, the FULL merged [nq, topk] global ids out on every rank (owned rows all-gathered)."""
        t = self.torch
        q = queries if isinstance(queries, t.Tensor) else t.from_numpy(np.ascontiguousarray(queries, np.float32))
        dev = t.device(device if device is not None else ("cuda" if t.cuda.is_available() else "cpu"))
        nq = q.shape[0]
        ctx, _ = self._stream_ctx(stream)
        with ctx:
            d_q = q.to(dev, non_blocking=True)
            ids, _ = self.batch_search_device(d_q, topk, ef, stream)
            if self.world == 1:
                out = ids
            else:
                rows = (nq + self.world - 1) // self.world
                mine = t.zeros((rows, topk), dtype=t.int32, device=dev)
                mine[: ids.shape[0]] = ids
                allr = t.empty((self.world, rows, topk), dtype=t.int32, device=dev)
                self.dist.all_gather_into_tensor(allr.view(self.world * rows, topk), mine, group=self.group)
                out = allr.permute(1, 0, 2).reshape(rows * self.world, topk)[:nq]  # row i = owner i % world, slot i // world
            res = out.cpu().numpy().view(np.uint32).copy()
        return res
