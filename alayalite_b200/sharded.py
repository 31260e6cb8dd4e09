"""ShardedIndex: ID-range sharding of one logical index over the GPUs of a box.

The reference has no multi-device mode; this is the one addition. One process per GPU
(torch.distributed, NCCL). Rank g owns the contiguous id range [offset_g, offset_g + n_g), builds its
own graph over it with local ids, and answers every query on its shard. The per-shard top-k lists
(fp32 distance, global id) are exchanged with one all-gather per batch and merged on the device by
(distance, id) -- exactly what a CPU merge of the reference's per-shard results would give.

torch is used for plumbing only (process group, device buffers); search and merge are this
package's CUDA kernels behind the C ABI.
"""

from __future__ import annotations

import numpy as np


class ShardedIndex:
    def __init__(self, local_index, id_offset: int, world: int, group=None, merge_fn=None, search_fn=None):
        """local_index: _cabi.NativeIndex of this rank's shard (already fitted).
        merge_fn / search_fn are injection points for the CPU (gloo) tests of the host logic."""
        import torch
        import torch.distributed as dist

        self.torch, self.dist = torch, dist
        self.ix, self.offset, self.world, self.group = local_index, int(id_offset), int(world), group
        self._merge = merge_fn or self._merge_device
        self._search = search_fn or self._search_device
        self._buf = {}

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

    def batch_search_device(self, d_queries, topk, ef, stream=None):
        """Queries already on this rank's device (replicated). Returns device tensors (ids, dist)."""
        nq = d_queries.shape[0]
        b = self._buffers(nq, topk, d_queries.device)
        self._search(d_queries, nq, topk, ef, b["ids"], b["dist"], stream)
        b["ids"] += self.offset  # local -> global ids
        return self.exchange(b["ids"], b["dist"], b, stream)

    def batch_search(self, queries, topk, ef, out=None, device=None, stream=None):
        """Host queries in (numpy or pinned torch tensor), merged global ids out on the host."""
        t = self.torch
        q = queries if isinstance(queries, t.Tensor) else t.from_numpy(np.ascontiguousarray(queries, np.float32))
        d_q = q.to(device if device is not None else ("cuda" if t.cuda.is_available() else "cpu"), non_blocking=True)
        ids, _ = self.batch_search_device(d_q, topk, ef, stream)
        if out is None:
            return ids.cpu().numpy().view(np.uint32)
        out.copy_(ids)
        if ids.is_cuda:
            t.cuda.current_stream().synchronize()
        return out
