// Shard exchange fused into the kernels that produce the final lists (traversal for raw spaces, re-rank for SQ).
//
// The reference has no multi-device mode. Here one logical index is split by contiguous id ranges
// over S GPUs; every shard answers every query of its replica group, and the S per-shard top-k lists of a query
// have to be folded into one by (distance, global id) -- the one real exchange step of the path. Instead of
// collective + barrier + merge kernel, the exchange rides on the search:
//
//   * query qi is OWNED by shard qi % S. A warp that finishes qi on shard s stores its list (ids + id offset,
//     distances, padding = (0xFFFFFFFF, FLT_MAX)) into slot s of the owner's mailbox -- plain stores into peer
//     memory mapped over NVLink (or local memory when s is the owner) --, fences at system scope and bumps the
//     owner's arrival counter of qi with a system-scope atomic;
//   * the warp whose increment completes the set (it observes S - 1) is the last shard to arrive: it reads the S
//     lists back from the owner's mailbox, ranks every entry by (distance, id) against every list (binary search
//     on the sorted lists) and writes the merged top-k into the owner's result rows, resets the arrival counter
//     and bumps the owner's `done` counter. No barrier, no merge kernel, nobody waits inside the search kernel;
//   * a rank that needs its owned rows on its own stream (one process per GPU) launches wait_done_kernel behind the
//     search: one thread polling `done` (with a time-out). In the single-process group the host simply
//     synchronises the S streams: every merge runs inside one of those kernels.
//
// Two parities of every buffer alternate between consecutive calls, so a fast shard may already work on call t+1
// while a slow one still merges call t (a shard cannot reach call t+2 before every shard has finished call t: its
// own call t+1 completes only after all lists of its owned queries have arrived).
#pragma once
#include <cfloat>

#include "common.cuh"

namespace ab {

constexpr int kMaxShards = 16;

// By value in the kernel arguments. Pointer i belongs to owner (shard) i of this replica group.
struct Exchange {
  int n_shards;           // 0 = no exchange
  uint32_t shard;         // this shard
  uint32_t id_offset;     // global id = local id + id_offset
  uint32_t rows;          // mailbox rows per slot (>= ceil(nq / n_shards))
  uint32_t compact;       // result row of query qi: qi / n_shards (owner-local) if set, else qi
  uint2 *lists[kMaxShards];               // [n_shards, rows, topk] {distance bits, global id}
  uint32_t *arrive[kMaxShards];           // [rows]
  unsigned long long *done[kMaxShards];   // [1]
  uint32_t *out_ids[kMaxShards];          // merged rows at the owner (or one shared host buffer)
  float *out_dist[kMaxShards];            // may be null
};

__device__ __forceinline__ uint2 ld_volatile_u2(const uint2 *p) {
  uint2 r;
  asm volatile("ld.volatile.global.v2.u32 {%0,%1}, [%2];" : "=r"(r.x), "=r"(r.y) : "l"(p) : "memory");
  return r;
}

// (distance, id) order of the merged list
__device__ __forceinline__ bool entry_less(float d0, uint32_t i0, float d1, uint32_t i1) { return d0 < d1 || (d0 == d1 && i0 < i1); }

// Warp-collective. loc_ids / loc_dist: this shard's final list of query qi (local ids), `nvalid` of its topk entries
// are real (the rest is the reference's zero padding and must not take part in the merge). scratch: shared memory
// of at least n_shards * topk * 8 bytes the warp may overwrite.
__device__ __forceinline__ void exchange_epilogue(const Exchange &x, uint32_t qi, uint32_t topk, uint32_t nvalid,
                                                  const uint32_t *loc_ids, const float *loc_dist, uint2 *scratch, int lane) {
  const uint32_t S = static_cast<uint32_t>(x.n_shards);
  const uint32_t o = qi % S, ql = qi / S;
  uint2 *box = x.lists[o] + size_t(ql) * topk;  // slot s of the owner: box + s * rows * topk
  uint2 *mine = box + size_t(x.shard) * x.rows * topk;
  for (uint32_t i = lane; i < topk; i += 32) {
    uint2 e = make_uint2(__float_as_uint(FLT_MAX), kEmptyId);
    if (i < nvalid) e = make_uint2(__float_as_uint(loc_dist[i]), loc_ids[i] + x.id_offset);
    mine[i] = e;
  }
  __threadfence_system();
  __syncwarp();
  uint32_t prev = 0;
  if (lane == 0) prev = atomicAdd_system(x.arrive[o] + ql, 1u);
  prev = __shfl_sync(0xFFFFFFFFu, prev, 0);
  if (prev != S - 1) return;

  // ---- last shard to arrive: merge ----------------------------------------------------------------
  __threadfence_system();
  const uint32_t total = S * topk;
  for (uint32_t t = lane; t < total; t += 32) {
    const uint32_t s = t / topk, i = t - s * topk;
    scratch[t] = ld_volatile_u2(box + size_t(s) * x.rows * topk + i);
  }
  __syncwarp();
  const uint32_t row = x.compact ? ql : qi;
  uint32_t *oi = x.out_ids[o] + size_t(row) * topk;
  float *od = x.out_dist[o] ? x.out_dist[o] + size_t(row) * topk : nullptr;
  uint32_t have = 0;
  for (uint32_t t = lane; t < total; t += 32) {
    const uint2 e = scratch[t];
    if (e.y == kEmptyId) continue;
    have++;
    const float d = __uint_as_float(e.x);
    uint32_t rank = 0;
    for (uint32_t s = 0; s < S; s++) {
      const uint2 *l = scratch + s * topk;
      // entries of list s strictly before e: all with a smaller distance (lists are sorted by distance) ...
      uint32_t lo = 0, hi = topk;
      while (lo < hi) {
        const uint32_t mid = (lo + hi) >> 1;
        if (__uint_as_float(l[mid].x) < d) lo = mid + 1; else hi = mid;
      }
      rank += lo;
      // ... plus those at the same distance with a smaller id
      for (uint32_t j = lo; j < topk && __uint_as_float(l[j].x) == d; j++) rank += (l[j].y < e.y) ? 1u : 0u;
    }
    if (rank < topk) {
      oi[rank] = e.y;
      if (od) od[rank] = d;
    }
  }
#pragma unroll
  for (int off = 16; off > 0; off >>= 1) have += __shfl_xor_sync(0xFFFFFFFFu, have, off);
  // slots no shard could fill stay 0 like the reference's zero-initialised pool
  for (uint32_t r = have + lane; r < topk; r += 32) {
    oi[r] = 0u;
    if (od) od[r] = 0.f;
  }
  __threadfence_system();
  __syncwarp();
  if (lane == 0) {
    x.arrive[o][ql] = 0u;  // nobody touches this counter again before the same parity comes round
    __threadfence_system();
    atomicAdd_system(x.done[o], 1ull);
  }
  __syncwarp();
}

}  // namespace ab
