"""ShardedIndex: ID-range sharding of one logical index over the GPUs of a box.

The reference has no multi-device mode; this is the one addition. One process per GPU
(torch.distributed, NCCL). Rank g owns the contiguous id range [offset_g, offset_g + n_g), builds its
own graph over it with local ids, and answers every query on its shard. The per-shard top-k lists
(fp32 distance, global id) are exchanged once per batch and merged on the device by (distance, id) --
exactly what a CPU merge of the reference's per-shard results would give.

Exchange. On one NVSwitch box the search kernel itself stores every final (global id, distance) into the
gather buffer of every rank -- symmetric memory mapped peer to peer -- so the transfer overlaps the search
and what remains per batch is one signal barrier and the merge kernel (3 launches, no NCCL call, no id-offset
pass). Without symmetric memory (CPU tests, other topologies) the lists go through all_gather_into_tensor.

torch is used for plumbing only (process group, device buffers); search and merge are this
package's CUDA kernels behind the C ABI.
"""

from __future__ import annotations

import numpy as np


class ShardedIndex:
    def __init__(self, local_index, id_offset: int, world: int, group=None, merge_fn=None, search_fn=None, peer_memory=True):
        """local_index: _cabi.NativeIndex of this rank's shard (already fitted).
        merge_fn / search_fn are injection points for the CPU (gloo) tests of the host logic."""
        import torch
        import torch.distributed as dist

        self.torch, self.dist = torch, dist
        self.ix, self.offset, self.world, self.group = local_index, int(id_offset), int(world), group
        self._merge = merge_fn or self._merge_device
        self._search = search_fn or self._search_device
        self._buf = {}
        self._peer = {}
        # peer-memory exchange needs the real kernels and more than one rank
        import os

        self.peer_memory = (bool(peer_memory) and search_fn is None and merge_fn is None and self.world > 1
                            and os.environ.get("AB_PEER_MEMORY", "1") != "0")
        self.exchange_kind = "none" if self.world == 1 else "nccl all_gather"
        self._step = 0

    # ---- the two device steps -------------------------------------------------------------------
    def _search_device(self, d_queries, nq, topk, ef, d_ids, d_dist, stream):
        self.ix.batch_search_device(d_queries.data_ptr(), nq, topk, ef, d_ids.data_ptr(), d_dist.data_ptr(), True, stream)

    def _merge_device(self, g_dist, g_ids, nq, topk, out_ids, out_dist, stream):
        from . import _cabi

        _cabi.merge_topk_device(g_dist.data_ptr(), g_ids.data_ptr(), self.world, nq, topk, out_ids.data_ptr(), out_dist.data_ptr(), stream)

    def _buffers(self, nq, topk, device):
        key = (nq, topk, str(device))
        if key not in self._buf:
            t = self.torch
            self._buf[key] = dict(
                ids=t.empty((nq, topk), dtype=t.int32, device=device), dist=t.empty((nq, topk), dtype=t.float32, device=device),
                g_ids=t.empty((self.world, nq, topk), dtype=t.int32, device=device), g_dist=t.empty((self.world, nq, topk), dtype=t.float32, device=device),
                m_ids=t.empty((nq, topk), dtype=t.int32, device=device), m_dist=t.empty((nq, topk), dtype=t.float32, device=device))
        return self._buf[key]

    def exchange(self, local_ids, local_dist, b, stream=None):
        """local (already global-id) lists -> all ranks hold the merged [nq, topk] result."""
        nq, topk = local_ids.shape
        if self.world == 1:
            return local_ids, local_dist
        # [world, nq, topk] filled rank by rank == concatenation along dim 0 of the [nq, topk] inputs
        self.dist.all_gather_into_tensor(b["g_ids"].view(self.world * nq, topk), local_ids, group=self.group)
        self.dist.all_gather_into_tensor(b["g_dist"].view(self.world * nq, topk), local_dist, group=self.group)
        self._merge(b["g_dist"], b["g_ids"], nq, topk, b["m_ids"], b["m_dist"], stream)
        return b["m_ids"], b["m_dist"]

    def _peer_buffers(self, nq, topk, device):
        """Symmetric [2 (parity), 2 (ids | dist), world, nq, topk] buffer + the peers' addresses of it, or None."""
        key = (nq, topk, str(device))
        if key not in self._peer:
            entry = None
            try:
                import torch.distributed._symmetric_memory as symm

                t, rank = self.torch, self.dist.get_rank(self.group)
                buf = symm.empty((2, 2, self.world, nq, topk), dtype=t.int32, device=device)
                hdl = symm.rendezvous(buf, self.group if self.group is not None else self.dist.group.WORLD)
                half = self.world * nq * topk * 4  # bytes of one [world, nq, topk] block
                entry = dict(buf=buf, hdl=hdl, rank=rank, ptrs=[int(p) for p in hdl.buffer_ptrs], half=half)
                self.exchange_kind = "peer stores from the search kernel + signal barrier"
            except Exception as e:  # noqa: BLE001  (no P2P / symmetric memory here: keep NCCL)
                self.peer_memory = False
                self.peer_error = repr(e)
            self._peer[key] = entry
        return self._peer[key]

    def _search_exchange_peer(self, d_queries, nq, topk, ef, b, pb, stream, after_search=None, out=None):
        t = self.torch
        parity = self._step & 1  # two buffer sets: a fast rank may already write step t+1 while a peer merges step t
        self._step += 1
        base = [p + parity * 2 * pb["half"] for p in pb["ptrs"]]
        self.ix.batch_search_scatter_device(d_queries.data_ptr(), nq, topk, ef, b["ids"].data_ptr(), b["dist"].data_ptr(), self.offset,
                                            pb["rank"], base, [p + pb["half"] for p in base], True, stream)
        if after_search is not None:
            after_search()
        # every rank's stores into this rank's buffer are complete and visible after the barrier (enqueued on the
        # stream the search ran on)
        if stream is None or isinstance(stream, t.cuda.Stream):
            with t.cuda.stream(stream):
                pb["hdl"].barrier(channel=0)
        else:
            with t.cuda.stream(t.cuda.ExternalStream(int(stream))):
                pb["hdl"].barrier(channel=0)
        g_ids = pb["buf"][parity, 0]
        g_dist = pb["buf"][parity, 1].view(t.float32)
        m_ids = b["m_ids"] if out is None else out  # a pinned host tensor is written by the merge kernel directly
        self._merge(g_dist, g_ids, nq, topk, m_ids, b["m_dist"], stream)
        return m_ids, b["m_dist"]

    def batch_search_device(self, d_queries, topk, ef, stream=None, after_search=None, device=None, out=None):
        """Queries on this rank's device (replicated), or in pinned host memory the device can read in place
        (then pass `device`). Returns device tensors (ids, dist); `out` (pinned host int32 [nq, topk]) receives the
        merged ids directly when the peer-memory exchange is in use.
        after_search: optional callable run right after the search launch (bench.py records an event there)."""
        nq = d_queries.shape[0]
        device = d_queries.device if device is None else self.torch.device(device)
        b = self._buffers(nq, topk, device)
        if self.peer_memory and device.type == "cuda":
            pb = self._peer_buffers(nq, topk, device)
            if pb is not None:
                return self._search_exchange_peer(d_queries, nq, topk, ef, b, pb, stream, after_search, out)
        self._search(d_queries, nq, topk, ef, b["ids"], b["dist"], stream)
        if after_search is not None:
            after_search()
        b["ids"] += self.offset  # local -> global ids
        return self.exchange(b["ids"], b["dist"], b, stream)

    def batch_search(self, queries, topk, ef, out=None, device=None, stream=None):
        """Host queries in (numpy or pinned torch tensor), merged global ids out on the host. Pinned tensors are
        read (queries) and written (out) in place by the kernels; anything else is copied."""
        t = self.torch
        q = queries if isinstance(queries, t.Tensor) else t.from_numpy(np.ascontiguousarray(queries, np.float32))
        dev = t.device(device if device is not None else ("cuda" if t.cuda.is_available() else "cpu"))
        in_place = dev.type == "cuda" and self.peer_memory and self._peer_buffers(q.shape[0], topk, dev) is not None
        d_q = q if (in_place and q.is_pinned()) else q.to(dev, non_blocking=True)
        direct_out = out if (in_place and out is not None and out.is_pinned()) else None
        ids, _ = self.batch_search_device(d_q, topk, ef, stream, device=dev, out=direct_out)
        if out is None:
            return ids.cpu().numpy().view(np.uint32)
        if direct_out is None:
            out.copy_(ids)
        if dev.type == "cuda":
            t.cuda.current_stream().synchronize()
        return out
