// Device code of the traversal kernel + its launcher, shared by the translation units that instantiate it
// (search.cu: raw rows; search_sq8.cu / search_sq4.cu: quantised spaces) so that they compile in parallel.
//
// Design. One warp owns one query at a time and pulls the next query index from a global counter (the role of the
// reference's task queue). Per hop the warp
//   1. holds the 128-byte adjacency row of the node being expanded (one coalesced load, lane j <- neighbour j),
//   2. filters the neighbours through the visited set: an exact table in shared memory (16-bit quotient entries in
//      two-choice buckets of four -- one 64-bit load per bucket, no probe sequence --, or one-entry slots with a
//      vote-driven probe loop, or 32-bit ids, whichever the id range allows); a query that outgrows it continues in the
//      warp's private overflow table in global memory (sized from ef, not from the number of rows) and, past that, in a
//      pooled N-bit bitmap (the reference's structure) -- never a false "visited" or "new",
//   3. evaluates all unvisited neighbours: sub-warps of TPV lanes per candidate row, 128-bit loads, three candidates
//      per sub-warp in flight, query chunks read from shared memory after the loads are issued, packed f32x2 math,
//      dequantisation fused for SQ codes (distance.cuh),
//   4. determines the next node to expand from the pool and the fresh distances and requests its adjacency row,
//   5. merges the accepted candidates into the sorted pool (LinearPool semantics: reject if full and d >= worst,
//      position = upper bound, cursor = first un-expanded entry) while that load is in flight.
// Distances are plain fp32; given equal distance values the traversal is step-for-step the reference's, so ids are
// identical up to fp32 summation-order effects (bit-identical on integer-valued data).
#pragma once
#include <cfloat>
#include <cstdlib>

#include "distance.cuh"
#include "search.cuh"

namespace ab {

namespace {

constexpr uint32_t kFull = 0xFFFFFFFFu;
constexpr int kR = 32;  // level-0 degree (reference: params_.max_nbrs_ is always 32 from Python)

struct WarpSmem {
  float *q_raw;       // raw query (raw space only)
  uint8_t *q_code;    // encoded query (SQ only)
  uint2 *pool;        // [ef + 1] entries {distance bits, id | expanded flag}
  uint32_t *hash;     // [hash_size]
  uint32_t *cand;     // [32] candidate ids of the current hop, compacted
  float *cdist;       // [32] their distances
};

// Visited set of one query: the shared-memory table (level 1), then the warp's overflow table in global memory
// (level 2), then a pooled N-bit bitmap (level 3). Derived values are recomputed from the kernel arguments where
// needed instead of living in registers across the hot loop.
struct Visited {
  uint32_t *hash;
  bool spilled;     // the shared-memory table is frozen, new ids go to the overflow table
  bool ovf_used;    // the warp's overflow table holds ids and must be cleaned after the query
  uint32_t listed;  // ids this query put into the overflow table (the first kSpillList entries are listed)
  int big;          // pooled bitmap this query claimed after filling the overflow table, or -1
};

// multiply-shift range reduction: any table size, no power-of-two rounding of the shared-memory footprint
__device__ __forceinline__ uint32_t hash_slot(uint32_t key, uint32_t size) { return __umulhi(key * 0x9E3779B1u, size); }

__device__ __forceinline__ uint32_t *ovf_slot(const SearchArgs &a) {
  return a.ovf + size_t(blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5)) * (size_t(a.ovf_size) + kSpillList);
}

// Level 3 of the visited set: claim one of the pooled N-bit bitmaps (rare: a query that fills its overflow table).
// Warps wait for a free one; holders always finish on their own, so this cannot deadlock.
// (Out-of-line helpers take the few values they need, not the argument block: a reference to the kernel parameters
// would force every thread to copy the whole block to local memory at kernel entry.)
__device__ __noinline__ int visit_claim_bitmap(uint32_t *big_flags, uint32_t n_big) {
  int got = 0;
  if ((threadIdx.x & 31) == 0) {
    uint32_t i = (blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5)) % n_big, tries = 0;
    while (atomicCAS(&big_flags[i], 0u, 1u) != 0u) {
      i = i + 1 == n_big ? 0u : i + 1;
      if (++tries >= n_big) {
        __nanosleep(2000);
        tries = 0;
      }
    }
    __threadfence();
    got = static_cast<int>(i);
  }
  return __shfl_sync(kFull, got, 0);
}

// Levels 2 and 3 of the visited set (warp-collective; `want` marks the lanes whose id was not settled by the
// shared-memory table). Level 2: the warp's private open-addressing table of ids in global memory; the probe loop
// is vote-driven like the shared-memory one. Level 3, entered when level 2 reaches 3/4 load: a pooled bitmap;
// level 2 is then frozen and only read. Returns true for ids seen for the first time.
__device__ __forceinline__ bool visit_overflow(const SearchArgs &a, Visited &vs, uint32_t key, bool want) {
  uint32_t *tab = ovf_slot(a);
  const uint32_t mask = a.ovf_size - 1;
  const int lane = threadIdx.x & 31;
  if (vs.big < 0 && vs.listed + 32 > a.ovf_limit) vs.big = visit_claim_bitmap(a.big_flags, a.n_big);
  bool is_new = false, pending = want, to_big = false;
  uint32_t slot = (key * 0x85EBCA6Bu) & mask, mine = 0;
  while (__any_sync(kFull, pending)) {
    if (pending) {
      uint32_t old;
      if (vs.big < 0) {
        old = atomicCAS(&tab[slot], kEmptyId, key);
        if (old == kEmptyId) {
          is_new = true;
          mine = slot;
        }
      } else {
        old = __ldcg(&tab[slot]);  // written by this warp's atomics earlier: read at L2, not through L1
        to_big = old == kEmptyId;
      }
      pending = !(old == kEmptyId || old == key);
      slot = (slot + 1) & mask;
    }
  }
  const uint32_t am = __ballot_sync(kFull, is_new);
  if (is_new) {  // remember the claimed entry for the clean-up
    const uint32_t pos = vs.listed + __popc(am & ((1u << lane) - 1));
    if (pos < kSpillList) tab[a.ovf_size + pos] = mine;
  }
  vs.listed += __popc(am);
  vs.ovf_used = true;
  if (vs.big >= 0) {
    if (to_big) {
      const uint32_t bit = 1u << (key & 31);
      is_new = (atomicOr(&a.big[size_t(vs.big) * a.big_words + (key >> 5)], bit) & bit) == 0;
    }
    __syncwarp();
  }
  return is_new;
}

// give the overflow table (and a claimed bitmap) back clean
__device__ __noinline__ void visit_overflow_release(uint32_t *tab, uint32_t ovf_size, uint32_t listed, int big_slot, uint32_t *big,
                                                    size_t big_words, uint32_t *big_flags) {
  const int lane = threadIdx.x & 31;
  __syncwarp();
  if (listed <= kSpillList) {
    for (uint32_t i = lane; i < listed; i += 32) tab[tab[ovf_size + i]] = kEmptyId;
  } else {
    uint4 *t4 = reinterpret_cast<uint4 *>(tab);
    for (uint32_t w = lane; w < ovf_size / 4; w += 32) t4[w] = make_uint4(kEmptyId, kEmptyId, kEmptyId, kEmptyId);
  }
  if (big_slot >= 0) {
    uint4 *bm = reinterpret_cast<uint4 *>(big + size_t(big_slot) * big_words);
    for (size_t w = lane; w < big_words / 4; w += 32) bm[w] = make_uint4(0, 0, 0, 0);
    __threadfence();
    __syncwarp();
    if (lane == 0) atomicExch(&big_flags[big_slot], 0u);
  }
  __syncwarp();
}

// bits 15 / 31 are CLEAR where the low / high halfword of p ^ r is zero (exact per halfword: no carry crosses the halves)
__device__ __forceinline__ uint32_t differs_halves(uint32_t p, uint32_t r) {
  const uint32_t t = p ^ r;
  return ((t & 0x7FFF7FFFu) + 0x7FFF7FFFu) | t;
}

// Level 1 of the visited set, two-choice bucket format (a.hash_buckets > 0). The table is hash_buckets buckets of four
// 16-bit entries, one 64-bit shared load each, filled from entry 0 upwards (0xFFFF = free). An id has two candidate
// buckets, taken from two different bijective scrambles of the b-bit id space (odd multipliers mod 2^b); an entry
// stores which of the two placed it (bit 15) and the remainder of THAT scramble inside its bucket (15 bits; at most
// 32767 ids of the id space per bucket, the plan checks), so an entry equal to (which, remainder) inside bucket B can
// only belong to one id: exact, like the slot format. A lookup reads both buckets and is a hit or a miss at once --
// no probe sequence, so the warp does not iterate as long as its unluckiest lane (the slot format at 3/4 load: a
// third of the kernel time at ef 280). A new id goes to the bucket with more free entries (ties: the first); with
// four entries per bucket that places 99.5 % of the ids at 75 % load and 96 % at 90 %. An id that finds both buckets
// full continues in the warp's overflow table (levels 2 / 3) -- and so will every later lookup of it, because
// buckets never lose entries during a query. The loop below only repeats for lanes whose CAS lost against another
// lane of the same warp (same bucket, same free entry).
__device__ __forceinline__ bool visit_buckets(const SearchArgs &a, Visited &vs, uint32_t key, bool active, bool &to_l2) {
  const uint32_t nb = a.hash_buckets;
  const uint32_t x1 = (key * 0x9E3779B1u) << a.hash_xshift, x2 = (key * 0x85EBCA6Bu) << a.hash_xshift;
  const uint32_t b1 = __umulhi(x1, nb), b2 = __umulhi(x2, nb);
  const uint32_t c1 = __umulhi(x1 * nb, a.hash_rmul), c2 = __umulhi(x2 * nb, a.hash_rmul) | 0x8000u;
  const uint32_t r1 = c1 * 0x10001u, r2 = c2 * 0x10001u;
  const volatile unsigned long long *tab = reinterpret_cast<const volatile unsigned long long *>(vs.hash);
  bool is_new = false, pending = active;
  to_l2 = false;
  while (__any_sync(kFull, pending)) {
    if (pending) {
      const unsigned long long B1 = tab[b1], B2 = tab[b2];
      const uint32_t p0 = static_cast<uint32_t>(B1), p1 = static_cast<uint32_t>(B1 >> 32);
      const uint32_t q0 = static_cast<uint32_t>(B2), q1 = static_cast<uint32_t>(B2 >> 32);
      const uint32_t all_differ = differs_halves(p0, r1) & differs_halves(p1, r1) & differs_halves(q0, r2) & differs_halves(q1, r2);
      if ((all_differ & 0x80008000u) != 0x80008000u) {
        pending = false;  // seen before
      } else {
        // free entries are a suffix of the bucket: 16 leading one-bits per free entry
        const int f1 = p1 == kFull ? 2 + (__clz(~p0) >> 4) : (__clz(~p1) >> 4);
        const int f2 = q1 == kFull ? 2 + (__clz(~q0) >> 4) : (__clz(~q1) >> 4);
        if ((f1 | f2) == 0) {
          pending = false;
          to_l2 = true;
        } else {
          const bool first = f1 >= f2;
          const int used = 4 - (first ? f1 : f2);  // the entry to claim
          const int wi = used >> 1, sh = (used & 1) << 4;
          const uint32_t old = first ? (wi ? p1 : p0) : (wi ? q1 : q0);
          const uint32_t want = old ^ ((0xFFFFu ^ (first ? c1 : c2)) << sh);
          if (atomicCAS(&vs.hash[(first ? b1 : b2) * 2u + wi], old, want) == old) {
            is_new = true;
            pending = false;
          }
        }
      }
    }
  }
  return is_new;
}

// Warp-collective: every lane with `active` tests-and-sets its key; true if the key was not yet in the set.
// The probe loop is driven by a warp vote, so the warp stays converged (a per-lane probe loop leaves the
// lanes split under independent thread scheduling, and every later warp-level primitive then takes the
// compiler's slow divergent path).
//
// Two table formats, both exact (never a false "visited", never a false "new"):
//  * 32-bit entries holding the id itself;
//  * 16-bit entries (a.hash_rbits > 0) when the id range is small enough relative to the table: the id is
//    scrambled by a bijection of the b-bit id space (odd multiplier mod 2^b), the table slot is the scaled
//    scramble, and the entry keeps what the slot does not already say -- the position inside the slot's
//    bucket of the id space (r, rbits bits) and the displacement from the home slot in probe steps. An
//    entry equal to (displacement, r) at slot home + displacement * step(r) can only belong to this id. Half the
//    shared memory per warp, which matters twice: more resident warps, and more of the unified L1/shared
//    array left to stage loads in flight. A probe that runs out of displacement range falls through to
//    the warp's overflow table (exact).
__device__ __forceinline__ bool visit(const SearchArgs &a, Visited &vs, uint32_t key, bool active) {
  const uint32_t hsize = a.hash_size;
  bool is_new = false;
  bool pending = active;
  bool to_bitmap = false;
  if (a.hash_buckets != 0) {
    is_new = visit_buckets(a, vs, key, active, to_bitmap);
  } else if (a.hash_rbits != 0) {
    const uint32_t x = (key * 0x9E3779B1u) << a.hash_xshift;
    uint32_t slot = __umulhi(x, hsize);
    uint32_t code = __umulhi(x * hsize, a.hash_rmul);  // displacement 0
    const uint32_t dstep = 1u << a.hash_rbits;
    // Probe sequence home, home + t, home + 2t, ... with an odd step t in 1..31 taken from the remainder: no primary
    // clustering (unsuccessful probes at 3/4 load: ~4 instead of ~8.5, and the warp iterates as long as its slowest
    // lane). Exactness is untouched: an entry (displacement d, remainder r) at slot s can only belong to the id with
    // home s - d * t(r) and remainder r.
    const uint32_t stride = 1u + 2u * (code & 15u);
    // Probe with plain reads; only an empty slot is claimed, with a 32-bit CAS on the word that holds it
    // (a failed CAS -- the slot or its neighbour half changed meanwhile -- re-reads the same slot).
    while (__any_sync(kFull, pending)) {
      if (pending) {
        const uint32_t cur = reinterpret_cast<const volatile uint16_t *>(vs.hash)[slot];  // one 16-bit shared load
        bool advance = true;
        if (cur == 0xFFFFu) {
          if (!vs.spilled) {
            advance = false;
            const uint32_t sh = (slot & 1u) << 4;
            const uint32_t w = vs.hash[slot >> 1];
            // (the half may have been claimed since the 16-bit read: then the same slot is simply probed again)
            if (((w >> sh) & 0xFFFFu) == 0xFFFFu && atomicCAS(&vs.hash[slot >> 1], w, w ^ ((0xFFFFu ^ code) << sh)) == w) {
              is_new = true;
              pending = false;
            }
          } else {
            to_bitmap = true;
            pending = false;
          }
        } else if (cur == code) {
          pending = false;
        }
        if (pending && advance) {
          slot += stride;  // double hashing: the step is a function of the remainder, which the entry stores
          slot = slot >= hsize ? slot - hsize : slot;
          code += dstep;
          if (code >= a.hash_dlimit) {  // out of displacement range
            pending = false;
            to_bitmap = true;
          }
        }
      }
    }
  } else {
    uint32_t slot = hash_slot(key, hsize);
    while (__any_sync(kFull, pending)) {
      if (pending) {
        uint32_t old;
        if (!vs.spilled) {
          old = atomicCAS(&vs.hash[slot], kEmptyId, key);
          is_new = old == kEmptyId;
        } else {
          old = vs.hash[slot];
          to_bitmap = old == kEmptyId;
        }
        pending = !(old == kEmptyId || old == key);
        slot = slot + 1 == hsize ? 0u : slot + 1;
      }
    }
  }
  if (__any_sync(kFull, to_bitmap)) {
    const bool r = visit_overflow(a, vs, key, to_bitmap);
    if (to_bitmap) is_new = r;
  }
  return is_new;
}

// LinearPool::insert (query_utils.hpp:254-272); warp-collective, all arguments warp-uniform.
// `expanded` inserts an entry that already carries the expanded flag (the hop that is about to run, see
// the main loop): it never becomes the cursor, and pushes the cursor one slot right when it lands at or
// before it.
__device__ __forceinline__ void pool_insert(uint2 *pool, int &size, int &cur, int ef, float d, uint32_t v, int lane,
                                            bool expanded = false) {
  int pos;  // upper bound: number of entries with dist <= d
  const int nsize = min(size + 1, ef);
  if (ef <= 64) {
    // the pool fits two entries per lane: read everything first, then every entry behind the insert
    // position moves one slot to the right in a single step
    const int i1 = lane + 32;
    const uint2 e0 = pool[lane];
    uint2 e1 = make_uint2(0u, 0u);
    if (size > 32 && i1 < size) e1 = pool[i1];
    pos = __popc(__ballot_sync(kFull, lane < size && __uint_as_float(e0.x) <= d));
    if (size > 32) pos += __popc(__ballot_sync(kFull, i1 < size && __uint_as_float(e1.x) <= d));
    __syncwarp();
    if (lane >= pos && lane < size && lane + 1 < nsize) pool[lane + 1] = e0;
    if (size > 32 && i1 >= pos && i1 < size && i1 + 1 < nsize) pool[i1 + 1] = e1;
    if (lane == 0) pool[pos] = make_uint2(__float_as_uint(d), v);
    __syncwarp();
  } else {
    pos = 0;
    for (int base = 0; base < size; base += 32) {
      const int i = base + lane;
      const uint32_t b = __ballot_sync(kFull, i < size && __uint_as_float(pool[i].x) <= d);
      pos += __popc(b);
      if (b != kFull) break;
    }
    for (int base = (nsize - 1) & ~31; base >= 0 && base + 31 > pos; base -= 32) {
      const int i = base + lane;
      const bool mv = i > pos && i < nsize;
      uint2 t = make_uint2(0u, 0u);
      if (mv) t = pool[i - 1];
      __syncwarp();
      if (mv) pool[i] = t;
      __syncwarp();
    }
    if (lane == 0) pool[pos] = make_uint2(__float_as_uint(d), v);
    __syncwarp();
  }
  size = nsize;
  if (expanded) {
    if (pos <= cur) cur = min(cur + 1, nsize);
  } else if (pos < cur) {
    cur = pos;
  }
}

// All accepted candidates of one hop at once (ef <= 64: two pool entries per lane). Equivalent to inserting them
// one by one in neighbour order: a candidate lands behind the pool entries with distance <= its own and behind
// the accepted candidates that precede it (smaller distance, or equal distance and earlier neighbour slot);
// a pool entry moves right by the number of candidates strictly closer than it; whatever ends at index >= ef is
// dropped (the sequential insert would have evicted or rejected it). `accept` / `d` / `v` are per lane, `acc` the
// ballot of `accept`, `sel` the lane whose candidate is the next hop (inserted with its expanded flag) or -1.
// The cursor follows the first un-expanded entry: the old cursor entry at its new index, or an inserted one.
__device__ __forceinline__ void pool_merge(uint2 *pool, const float *cdist, uint32_t newmask, int &size, int &cur, int ef,
                                           uint32_t acc, int sel, bool accept, float d, uint32_t v, int lane) {
  const bool two = size > 32;  // warp-uniform
  const bool valid0 = lane < size, valid1 = two && lane + 32 < size;
  uint2 e0 = make_uint2(0u, 0u), e1 = make_uint2(0u, 0u);
  if (valid0) e0 = pool[lane];
  if (valid1) e1 = pool[lane + 32];
  const float d0 = __uint_as_float(e0.x), d1 = __uint_as_float(e1.x);
  int shift0 = 0, shift1 = 0, mypos = 0, before = 0;
  const int nacc = __popc(acc);
  for (uint32_t rest = acc; rest;) {
    const int j = __ffs(rest) - 1;
    rest &= rest - 1;
    const float dj = cdist[__popc(newmask & ((1u << j) - 1))];  // uniform shared-memory read
    shift0 += (valid0 && dj < d0) ? 1 : 0;
    int le = __popc(__ballot_sync(kFull, valid0 && d0 <= dj));
    if (two) {
      shift1 += (valid1 && dj < d1) ? 1 : 0;
      le += __popc(__ballot_sync(kFull, valid1 && d1 <= dj));
    }
    if (lane == j) mypos = le;
    before += (accept && lane != j && (dj < d || (dj == d && j < lane))) ? 1 : 0;
  }
  mypos += before;
  const int n0 = lane + shift0, n1 = lane + 32 + shift1;
  // cursor: new index of the entry it pointed at (if it survives), else the end; then any un-expanded insert before it
  int c = size;
  if (cur < size) {
    const int sh = __shfl_sync(kFull, cur >= 32 ? shift1 : shift0, cur & 31);
    c = cur + sh;
  } else {
    c = size + nacc;
  }
  const int nsize = min(size + nacc, ef);
  const int ins = (accept && lane != sel && mypos < ef) ? mypos : 0x7FFFFFFF;
  c = min(min(c, nsize), static_cast<int>(__reduce_min_sync(kFull, static_cast<unsigned>(ins))));
  __syncwarp();
  if (valid0 && n0 < ef) pool[n0] = e0;
  if (valid1 && n1 < ef) pool[n1] = e1;
  if (accept && mypos < ef) pool[mypos] = make_uint2(__float_as_uint(d), lane == sel ? (v | kCheckedBit) : v);
  __syncwarp();
  size = nsize;
  cur = c;
}

// The same merge for any pool size (used above ef = 64). Every accepting lane finds its candidate's upper bound p
// in the sorted pool by binary search (all lanes in parallel, log2(size) + 1 uniform steps). With
//   before_j = accepted candidates ordered before j (smaller distance, or equal distance and earlier slot)
// candidate j lands at f_j = p_j + before_j; these final positions are distinct, and the old entries fill the
// remaining positions in their old order. So the new pool is produced by DESTINATION: for position f that no
// candidate takes, the source is old entry f - #{j : f_j < f}. Per 32-position tile that is one redux.or (which
// positions of the tile go to candidates), one vote (how many candidates land below the tile) and a popc per lane,
// independent of the number of accepted candidates. Tiles are rewritten from the top down to the tile of the
// smallest f_j: a tile reads sources at or below its own positions, and everything above it is already final, so
// no write can land on an entry that is still to be read; the candidates are stored last. Whatever would end at
// index >= ef is never copied. One pass replaces up to 32 sequential LinearPool::insert calls.
__device__ __forceinline__ void pool_merge_tiles(uint2 *pool, int &size, int &cur, int ef, uint32_t acc, int sel, bool accept,
                                                 float d, uint32_t v, int lane) {
  // 1. upper bound of every accepted candidate
  int p = 0;
  for (int b = 1 << (31 - __clz(size | 1)); b > 0; b >>= 1) {
    const int t = p + b;
    if (accept && t <= size && __uint_as_float(pool[t - 1].x) <= d) p = t;
  }
  const int nacc = __popc(acc);
  // 2. rank among the accepted candidates
  int before = 0;
  for (uint32_t rest = acc; rest;) {
    const int j = __ffs(rest) - 1;
    rest &= rest - 1;
    const float dj = __shfl_sync(kFull, d, j);
    before += (accept && lane != j && (dj < d || (dj == d && j < lane))) ? 1 : 0;
  }
  const int mypos = accept ? p + before : 0x7FFFFFFF;
  // cursor: new index of the entry it pointed at, or the first un-expanded insert before it
  int c;
  if (cur < size) {
    c = cur + __popc(__ballot_sync(kFull, accept && p <= cur));
  } else {
    c = size + nacc;
  }
  const int nsize = min(size + nacc, ef);
  const int ins = (accept && lane != sel && mypos < ef) ? mypos : 0x7FFFFFFF;
  c = min(min(c, nsize), static_cast<int>(__reduce_min_sync(kFull, static_cast<unsigned>(ins))));
  // 3. rewrite positions fmin .. nsize - 1, top tile first
  const int fmin = static_cast<int>(__reduce_min_sync(kFull, static_cast<unsigned>(mypos)));
  for (int base = (nsize - 1) & ~31; base >= 0 && base + 31 >= fmin; base -= 32) {
    const uint32_t taken = __reduce_or_sync(kFull, (mypos >> 5) == (base >> 5) ? (1u << (mypos & 31)) : 0u);
    const int below = __popc(__ballot_sync(kFull, mypos < base));
    const int f = base + lane;
    const bool mv = f < nsize && f >= fmin && ((taken >> lane) & 1u) == 0;
    uint2 e = make_uint2(0u, 0u);
    if (mv) e = pool[f - below - __popc(taken & ((1u << lane) - 1))];
    __syncwarp();
    if (mv) pool[f] = e;
  }
  __syncwarp();
  if (accept && mypos < ef) pool[mypos] = make_uint2(__float_as_uint(d), lane == sel ? (v | kCheckedBit) : v);
  __syncwarp();
  size = nsize;
  cur = c;
}

// order-preserving map float -> uint32 (for warp-wide integer min)
__device__ __forceinline__ uint32_t float_key(float f) {
  const uint32_t b = __float_as_uint(f);
  return b ^ ((b >> 31) ? 0xFFFFFFFFu : 0x80000000u);
}

// Evaluate the search-space distance of the candidate held by every lane with want == true.
// Generic path: query in shared memory, any space (distance.cuh row_distance).
template <int QUANT, bool IP, int TPV, int MAXC>
__device__ __forceinline__ float eval_lanes(const SearchArgs &a, const WarpSmem &ws, const QuerySmem &qs,
                                            uint32_t v, bool want, int lane) {
  constexpr int G = 32 / TPV;
  const uint32_t mask = __ballot_sync(kFull, want);
  const int cnt = __popc(mask);
  const int rank = __popc(mask & ((1u << lane) - 1));
  if (want) ws.cand[rank] = v;
  __syncwarp();
  const int g = lane / TPV, lg = lane % TPV;
  for (int base = 0; base < cnt; base += G) {
    const int k = base + g;
    const bool act = k < cnt;
    const uint32_t cv = act ? ws.cand[k] : 0u;
    float d = row_distance<QUANT, IP, TPV, MAXC>(a.rows + size_t(cv) * size_t(a.row_stride), a.nchunks, lg, qs, act);
    if (act && lg == 0) ws.cdist[k] = d;
  }
  __syncwarp();
  const float out = want ? ws.cdist[rank] : FLT_MAX;
  __syncwarp();
  return out;
}

// Raw-space fast path: query slice in registers, packed fp32 math, two candidates per group in
// flight (distance.cuh raw_distance2).
template <bool IP, int TPV, int CPL, bool EXACT>
__device__ __forceinline__ float eval_lanes_raw(const SearchArgs &a, const WarpSmem &ws, const U64x2 (&q)[CPL], uint32_t v,
                                                bool want, int lane) {
  constexpr int G = 32 / TPV;
  const uint32_t mask = __ballot_sync(kFull, want);
  const int cnt = __popc(mask);
  const int rank = __popc(mask & ((1u << lane) - 1));
  if (want) ws.cand[rank] = v;
  __syncwarp();
  const int g = lane / TPV, lg = lane % TPV;
  const uint8_t *lane_base = a.rows + lg * 16;
  for (int base = 0; base < cnt; base += 2 * G) {
    const int k0 = base + g, k1 = k0 + G;
    const bool act0 = k0 < cnt, act1 = k1 < cnt;
    const uint32_t c0 = ws.cand[act0 ? k0 : 0];
    float d0;
    uint32_t c1 = c0;
    if (cnt - base > G) {  // warp-uniform: a second candidate exists for at least one group
      c1 = ws.cand[act1 ? k1 : 0];
      raw_distance2<IP, TPV, CPL, EXACT>(lane_base + size_t(c0) * a.row_stride, lane_base + size_t(c1) * a.row_stride, q,
                                         a.nchunks, lg, d0);
    } else {  // every lane of the group gets the sum
      d0 = raw_distance1<IP, TPV, CPL, EXACT>(lane_base + size_t(c0) * a.row_stride, q, a.nchunks, lg);
    }
    // lane 0 of a group ends up with the first candidate's distance, lane 1 with the second's
    if (lg < 2) {
      const int k = lg ? k1 : k0;
      if (k < cnt) ws.cdist[k] = d0;
    }
  }
  __syncwarp();
  const float out = want ? ws.cdist[rank] : FLT_MAX;
  __syncwarp();
  return out;
}

// Three candidates per group in flight, query read from shared memory after the loads are issued
// (distance.cuh raw_distance_multi / sq_distance_multi). Used when a row is at most 4 chunks per lane.
template <int QUANT, bool IP, int TPV, int CPL, bool EXACT, int NC>
__device__ __forceinline__ float multi_distance(const uint8_t *r0, const uint8_t *r1, const uint8_t *r2, const QuerySmem &qs,
                                                int nchunks, int lg) {
  if constexpr (QUANT == AB_QUANT_NONE) {
    return raw_distance_multi<IP, TPV, CPL, EXACT, NC>(r0, r1, r2, qs.q_raw, nchunks, lg);
  } else {
    return sq_distance_multi<QUANT, IP, TPV, CPL, NC>(r0, r1, r2, qs, nchunks, lg);
  }
}

// candidates per lane group kept in flight by the multi-candidate path, and resident CTAs the raw kernels are
// built for (compile-time experiment knobs, see scripts/build_variant.sh)
#ifndef AB_INFLIGHT
#define AB_INFLIGHT 3
#endif
#ifndef AB_MINBLOCKS
#define AB_MINBLOCKS 3
#endif
template <int QUANT, bool IP, int TPV, int CPL, bool EXACT, int INFL>
__device__ __forceinline__ float eval_lanes_multi(const SearchArgs &a, const WarpSmem &ws, const QuerySmem &qs, uint32_t v,
                                                  bool want, int lane) {
  constexpr int G = 32 / TPV;
  const uint32_t mask = __ballot_sync(kFull, want);
  const int cnt = __popc(mask);
  const int rank = __popc(mask & ((1u << lane) - 1));
  if (want) ws.cand[rank] = v;
  __syncwarp();
  const int g = lane / TPV, lg = lane % TPV;
  const uint8_t *lane_base = a.rows + lg * 16;
  if ((a.prefetch & 1) && cnt > INFL * G) {
    // more than one pass: the rows of the later passes start their way to L2 now, one bulk request per row
    const int k = INFL * G + lane;
    if (k < cnt) bulk_prefetch_l2(a.rows + size_t(ws.cand[k]) * a.row_stride, static_cast<uint32_t>(a.nchunks) * 16u);
  }
  for (int base = 0; base < cnt; base += INFL * G) {
    const int rem = cnt - base;  // warp-uniform
    const int k0 = base + g, k1 = k0 + G, k2 = k1 + G;
    const uint8_t *r0 = lane_base + size_t(ws.cand[k0 < cnt ? k0 : 0]) * a.row_stride;
    if constexpr (INFL == 6 && QUANT == AB_QUANT_NONE) {
      if (rem > 3 * G) {  // deep pass: four to six candidates per group
        const uint8_t *r[6];
        r[0] = r0;
#pragma unroll
        for (int c = 1; c < 6; c++) {
          const int kc = k0 + c * G;
          r[c] = lane_base + size_t(ws.cand[kc < cnt ? kc : 0]) * a.row_stride;
        }
        float d_hi, d_lo;
        if (rem > 5 * G) {
          d_lo = raw_distance_deep<IP, TPV, CPL, EXACT, 6>(r, qs.q_raw, a.nchunks, lg, d_hi);
        } else if (rem > 4 * G) {
          d_lo = raw_distance_deep<IP, TPV, CPL, EXACT, 5>(r, qs.q_raw, a.nchunks, lg, d_hi);
        } else {
          d_lo = raw_distance_deep<IP, TPV, CPL, EXACT, 4>(r, qs.q_raw, a.nchunks, lg, d_hi);
        }
        // lanes 0..2 of the group hold candidates 0..2 (d_lo) and 3..5 (d_hi)
        const int klo = k0 + lg * G, khi = klo + 3 * G;
        if (lg < 3 && klo < cnt) ws.cdist[klo] = d_lo;
        if (lg < 3 && khi < cnt) ws.cdist[khi] = d_hi;
        continue;
      }
    }
    if (INFL > 2 && rem > 2 * G) {
      const uint8_t *r1 = lane_base + size_t(ws.cand[k1 < cnt ? k1 : 0]) * a.row_stride;
      const uint8_t *r2 = lane_base + size_t(ws.cand[k2 < cnt ? k2 : 0]) * a.row_stride;
      const float d = multi_distance<QUANT, IP, TPV, CPL, EXACT, 3>(r0, r1, r2, qs, a.nchunks, lg);
      const int k = k0 + lg * G;  // lanes 0..2 of the group hold candidates 0..2
      if (lg < 3 && k < cnt) ws.cdist[k] = d;
    } else if (rem > G) {
      const uint8_t *r1 = lane_base + size_t(ws.cand[k1 < cnt ? k1 : 0]) * a.row_stride;
      const float d = multi_distance<QUANT, IP, TPV, CPL, EXACT, 2>(r0, r1, r1, qs, a.nchunks, lg);
      const int k = k0 + lg * G;
      if (lg < 2 && k < cnt) ws.cdist[k] = d;
    } else {
      const float d = multi_distance<QUANT, IP, TPV, CPL, EXACT, 1>(r0, r0, r0, qs, a.nchunks, lg);
      if (lg == 0 && k0 < cnt) ws.cdist[k0] = d;
    }
  }
  __syncwarp();
  const float out = want ? ws.cdist[rank] : FLT_MAX;
  __syncwarp();
  return out;
}

// CPL > 0 selects the raw-space fast path (needs QUANT == none and TPV * CPL >= nchunks; EXACT: ==);
// CPL == 0 is the generic path for any space.
// INFL: candidates per lane group in flight (3, or 6 = the deep variant: 12 / 16 resident warps by registers)
template <int QUANT, bool IP, int TPV, int CPL, bool EXACT, int INFL = AB_INFLIGHT>
__global__ void __launch_bounds__((INFL == 6 && CPL == 4) ? 192 : 256,
                                  INFL == 6 ? 2 : ((QUANT == AB_QUANT_NONE && CPL > 0 && CPL <= 4) ? AB_MINBLOCKS : 2))
    beam_search_kernel(const SearchArgs a) {
  constexpr bool FAST = QUANT == AB_QUANT_NONE && CPL > 0;
  constexpr int MAXC = 8;
  // three candidates in flight, query from shared memory (raw rows, and SQ codes with packed arithmetic)
  constexpr bool FAST3 = CPL > 0 && CPL <= 4;
  constexpr int QREGS = (FAST && !FAST3) ? CPL : 1;
  extern __shared__ __align__(16) uint8_t smem[];
  const int lane = lane_id();
  // redux lands in a uniform register: the compiler then knows every per-warp address below is warp-uniform
  const int warp = static_cast<int>(__reduce_min_sync(kFull, threadIdx.x >> 5));
  const int warps = blockDim.x >> 5;

  // ---- carve shared memory: CTA-wide tables, then one region per warp -------------------------
  float *s_scale = reinterpret_cast<float *>(smem);
  float *s_vmin = s_scale + (QUANT != AB_QUANT_NONE ? a.table_len : 0);
  size_t off = (QUANT != AB_QUANT_NONE ? size_t(a.table_len) * 8 : 0);
  const size_t q_bytes = size_t(a.nchunks) * 16;
  const size_t pool_bytes = round_up16(size_t(a.ef + 1) * 8);
  const size_t per_warp = 256 + q_bytes + pool_bytes + size_t(a.hash_words) * 4;
  uint8_t *wbase = smem + off + size_t(warp) * per_warp;
  WarpSmem ws;
  ws.cand = reinterpret_cast<uint32_t *>(wbase);
  ws.cdist = reinterpret_cast<float *>(wbase + 128);
  ws.q_raw = reinterpret_cast<float *>(wbase + 256);
  ws.q_code = wbase + 256;
  ws.pool = reinterpret_cast<uint2 *>(wbase + 256 + q_bytes);
  ws.hash = reinterpret_cast<uint32_t *>(wbase + 256 + q_bytes + pool_bytes);

  if (QUANT != AB_QUANT_NONE) {
    // The multi-candidate path reads the tables of chunk (lg + i * TPV): consecutive lanes, consecutive chunks.
    // Stored in dimension order that is a 64 / 128-byte stride -> every 128-bit read of a quarter warp lands in
    // the same banks. Transposed to [float4 index within the chunk][chunk] the same reads are conflict-free.
    constexpr int DPC = QUANT == AB_QUANT_SQ8 ? 16 : 32;  // dimensions per 16-byte chunk of codes
    for (int i = threadIdx.x; i < a.table_len; i += blockDim.x) {
      int at = i;
      if (CPL > 0) {
        const int chunk = i / DPC, e = i % DPC;
        at = ((e >> 2) * a.nchunks + chunk) * 4 + (e & 3);
      }
      if (CPL == 0 || i / DPC < a.nchunks) {
        // packed SQ4 path: high nibbles are decoded as 16 x code (distance.cuh word_pairs), their scale is s / 16
        s_scale[at] = (QUANT == AB_QUANT_SQ4 && CPL > 0 && (i & 1)) ? a.scale[i] * 0.0625f : a.scale[i];
        s_vmin[at] = a.vmin[i];
      }
    }
    __syncthreads();
  }
  QuerySmem qs;
  qs.q_raw = ws.q_raw;
  qs.q_code = ws.q_code;
  qs.scale = s_scale;
  qs.vmin = s_vmin;

  Visited vs;
  vs.hash = ws.hash;
  const int ef = static_cast<int>(a.ef);

  while (true) {
    if (lane == 0) ws.cand[0] = atomicAdd(a.work_counter, 1u);
    __syncwarp();
    const uint32_t qi = ws.cand[0];
    __syncwarp();
    if (qi >= a.nq) break;
    const float *qg = a.queries + size_t(a.q_ids ? a.q_ids[qi] : qi) * a.q_stride;
    if (a.q_ready != nullptr) {
      // the query batch is still being copied in (search_host, pageable caller buffers): wait for this query's chunk
      if (lane == 0) {
        const volatile uint32_t *flag = a.q_ready + qi / a.q_chunk;
        const long long t0 = clock64();
        while (*flag == 0u && clock64() - t0 < 4000000000ll) __nanosleep(100);  // the time-out only guards against a failed copy
        __threadfence();
      }
      __syncwarp();
    }

    // ---- stage the query -------------------------------------------------------------------
    if (QUANT == AB_QUANT_NONE) {
      // RawSpace::QueryComputer (raw_space.hpp:264-274): aligned copy, COS normalises first
      const int nfl = a.nchunks * 4;
      float ss = 0.f;
      for (int i = lane; i < nfl; i += 32) {
        const float x = i < static_cast<int>(a.dim) ? __ldcg(qg + i) : 0.f;  // read at L2: the batch may still be arriving
        ws.q_raw[i] = x;
        ss = fmaf(x, x, ss);
      }
      if (a.metric == AB_METRIC_COS && !a.q_prenormalized) {
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) ss += __shfl_xor_sync(kFull, ss, o);
        const float inv = static_cast<float>(1.0 / static_cast<double>(sqrtf(ss)));  // data_utils.hpp:36-46
        __syncwarp();
        for (int i = lane; i < nfl; i += 32) ws.q_raw[i] *= inv;
      }
    } else if (QUANT == AB_QUANT_SQ8) {
      // SQ8Space::QueryComputer (sq8_space.hpp:264-269): encode the (un-normalised) query
      const int nb = a.nchunks * 16;
      for (int i = lane; i < nb; i += 32) {
        uint32_t c = 0;
        if (i < static_cast<int>(a.dim)) c = sq_quantize(__ldcg(qg + i), a.qmin[i], a.qmax[i], 255.f);
        ws.q_code[i] = static_cast<uint8_t>(c);
      }
    } else {
      // SQ4Space::QueryComputer (sq4_space.hpp:265-270) + SQ4Quantizer::encode (sq4.hpp:139-149)
      const int nb = a.nchunks * 16;
      for (int j = lane; j < nb; j += 32) {
        uint32_t first = 0, second = 0;
        const int i = 2 * j;
        if (i < static_cast<int>(a.dim)) first = sq_quantize(__ldcg(qg + i), a.qmin[i], a.qmax[i], 15.f);
        if (i + 1 < static_cast<int>(a.dim)) second = sq_quantize(__ldcg(qg + i + 1), a.qmin[i + 1], a.qmax[i + 1], 15.f);
        ws.q_code[j] = static_cast<uint8_t>((first << 4) | second);
      }
    }
    for (uint32_t i = lane; i < a.hash_words / 4; i += 32) reinterpret_cast<uint4 *>(ws.hash)[i] = make_uint4(kEmptyId, kEmptyId, kEmptyId, kEmptyId);
    for (int i = lane; i <= ef; i += 32) {
      ws.pool[i] = make_uint2(0u, 0u);
    }
    __syncwarp();

    int size = 0, cur = 0;
    uint32_t nvis = 0;
    uint32_t ndc = 0, hops = 0, ndc_ov = 0;  // per query: far below 2^32
    vs.spilled = false;
    vs.ovf_used = false;
    vs.listed = 0;
    vs.big = -1;

    // ---- entry point: OverlayGraph::initialize (overlay_graph.hpp:122-144) ---------------------
    uint32_t u = a.ep;
    U64x2 qreg[QREGS];
    if (FAST && !FAST3) load_query_regs<TPV, QREGS, EXACT>(ws.q_raw, a.nchunks, lane % TPV, qreg);
    // distance of the candidate in every lane with WANT; afterwards ws.cand / ws.cdist hold the wanted ids and
    // distances compacted in lane order. Warp-uniform values are re-read from there (uniform shared-memory
    // loads) rather than shuffled, which keeps the control flow provably convergent for the compiler.
    auto eval = [&](uint32_t id, bool want) -> float {
      float dd;
      if constexpr (FAST3) {
        dd = eval_lanes_multi<QUANT, IP, TPV, CPL, EXACT, INFL>(a, ws, qs, id, want, lane);
      } else if constexpr (FAST) {
        dd = eval_lanes_raw<IP, TPV, QREGS, EXACT>(a, ws, qreg, id, want, lane);
      } else {
        dd = eval_lanes<QUANT, IP, TPV, MAXC>(a, ws, qs, id, want, lane);
      }
      if (QUANT == AB_QUANT_NONE && a.valid != nullptr) {
        // RawSpace::QueryComputer returns FLT_MAX for tombstoned rows (raw_space.hpp:297-304)
        const uint32_t wm = __ballot_sync(kFull, want);
        if (want && ((a.valid[id >> 5] >> (id & 31)) & 1u) == 0) {
          dd = FLT_MAX;
          ws.cdist[__popc(wm & ((1u << lane) - 1))] = dd;
        }
        __syncwarp();
      }
      return dd;
    };
#define AB_EVAL(ID, WANT) eval(ID, WANT)
    AB_EVAL(u, lane == 0);
    float cur_d = ws.cdist[0];
    __syncwarp();
    ndc_ov += 1;
    if (a.ov_off != nullptr) {
      for (int level = static_cast<int>(a.ep_level); level > 0; --level) {
        bool changed = true;
        while (changed) {
          changed = false;
          const uint32_t row = a.ov_off[u] + static_cast<uint32_t>(level - 1);
          const uint32_t v = a.ov_lists[size_t(row) * kR + lane];
          const uint32_t empties = __ballot_sync(kFull, v == kEmptyId);
          const int cnt = empties ? __ffs(empties) - 1 : kR;
          const bool want = lane < cnt;
          const float d = AB_EVAL(v, want);
          ndc_ov += cnt;
          // sequential "if (dist < cur_dist)" scan == first index attaining the minimum, if < cur_d
          if (cnt > 0) {
            const uint32_t key = want ? float_key(d) : 0xFFFFFFFFu;
            const uint32_t kmin = __reduce_min_sync(kFull, key);
            const int bl = __ffs(__ballot_sync(kFull, want && key == kmin)) - 1;  // want lanes are 0..cnt-1: rank == lane
            const float best = ws.cdist[bl];
            if (best < cur_d) {
              cur_d = best;
              u = ws.cand[bl];
              changed = true;
            }
          }
          __syncwarp();
        }
      }
    }
    if (lane == 0) atomicAdd(&a.stats[2], static_cast<unsigned long long>(ndc_ov));
    pool_insert(ws.pool, size, cur, ef, cur_d, u, lane);
    visit(a, vs, u, lane == 0);
    nvis = 1;
    __syncwarp();

    // ---- best-first search on level 0 (graph_search_job.hpp:307-331) ---------------------------
    // Same sequence of pops and inserts as the reference, scheduled differently: once the distances of a
    // hop are known, the next node to expand is already determined -- the closer of the first un-expanded
    // pool entry and the best accepted candidate (ties go to the pool entry, which upper-bound insertion
    // would have placed first) -- so its adjacency row is requested first and the pool inserts of this hop
    // run while that load is in flight. A candidate chosen this way is inserted with its expanded flag set.
    u = ws.pool[0].y;  // LinearPool::pop of the entry node
    __syncwarp();
    if (lane == 0) ws.pool[0].y = u | kCheckedBit;
    cur = 1;
    uint32_t v = a.adj[size_t(u) * kR + lane];
    while (true) {
      hops++;
      const uint32_t empties = __ballot_sync(kFull, v == kEmptyId);
      const int cnt = empties ? __ffs(empties) - 1 : kR;
      if ((a.prefetch & 2) && lane < cnt) bulk_prefetch_l2(a.rows + size_t(v) * a.row_stride, static_cast<uint32_t>(a.nchunks) * 16u);
      if (!vs.spilled && nvis + kR > a.hash_limit) vs.spilled = true;
      const bool is_new = visit(a, vs, v, lane < cnt);
      const uint32_t newmask = __ballot_sync(kFull, is_new);
      float d = FLT_MAX;
      if (newmask) {
        nvis += __popc(newmask);
        d = AB_EVAL(v, is_new);
      }
      // A candidate that fails "d < worst" against a full pool now can only fail later too (worst never
      // grows), so pre-filtering is exact.
      const float worst = size == ef ? __uint_as_float(ws.pool[size - 1].x) : FLT_MAX;
      const bool accept = is_new && (size < ef || d < worst);
      uint32_t acc = __ballot_sync(kFull, accept);
      const bool have_old = cur < size;
      if (acc == 0 && !have_old) break;
      int sel = -1;
      if (acc) {
        const uint32_t key = accept ? float_key(d) : 0xFFFFFFFFu;
        const uint32_t kmin = __reduce_min_sync(kFull, key);
        const int bl = __ffs(__ballot_sync(kFull, accept && key == kmin)) - 1;  // first neighbour attaining the minimum
        const int rb = __popc(newmask & ((1u << bl) - 1));
        if (!have_old || ws.cdist[rb] < __uint_as_float(ws.pool[cur].x)) {
          sel = bl;
          u = ws.cand[rb];
        }
      }
      if (sel < 0) {
        // LinearPool::pop (query_utils.hpp:284-293)
        u = ws.pool[cur].y;
        __syncwarp();
        if (lane == 0) ws.pool[cur].y = u | kCheckedBit;
        __syncwarp();
        int nxt = size;
        for (int base = cur + 1; base < size; base += 32) {
          const int i = base + lane;
          const uint32_t b = __ballot_sync(kFull, i < size && (ws.pool[i].y & kCheckedBit) == 0);
          if (b) {
            nxt = base + __ffs(b) - 1;
            break;
          }
        }
        cur = nxt;
      }
      const uint32_t v_next = a.adj[size_t(u) * kR + lane];

      // inserts in neighbour order, all accepted candidates of the hop in one merge step
      if (acc) {
        if (ef <= 64) {
          pool_merge(ws.pool, ws.cdist, newmask, size, cur, ef, acc, sel, accept, d, v, lane);
        } else {
          pool_merge_tiles(ws.pool, size, cur, ef, acc, sel, accept, d, v, lane);
        }
      }
      // the closest un-expanded entry is the likely hop after this one: pull its adjacency row towards L2
      if (lane == 0 && cur < size) {
        const uint32_t nx = ws.pool[cur].y & kIdMask;
        asm volatile("prefetch.global.L2 [%0];" ::"l"(a.adj + size_t(nx) * kR));
      }
      v = v_next;
    }
    ndc = nvis - 1;

    // ---- results ----------------------------------------------------------------------------
    for (uint32_t i = lane; i < a.topk; i += 32) {
      const bool in = i <= a.ef;  // slots past the pool size are zero, like the reference's pool
      uint32_t rid = in ? (ws.pool[i].y & kIdMask) : 0u;
      float rd = in ? __uint_as_float(ws.pool[i].x) : 0.f;
      if (a.pad_empty && static_cast<int>(i) >= size) {
        rid = kEmptyId;
        rd = FLT_MAX;
      }
      a.out_ids[size_t(qi) * a.topk + i] = rid;
      if (a.out_dist) a.out_dist[size_t(qi) * a.topk + i] = rd;
    }
    if (a.pool_ids) {
      for (int i = lane; i < ef; i += 32) {
        a.pool_ids[size_t(qi) * a.ef + i] = i < size ? (ws.pool[i].y & kIdMask) : 0u;
        a.pool_dist[size_t(qi) * a.ef + i] = i < size ? __uint_as_float(ws.pool[i].x) : 0.f;
      }
      if (lane == 0) a.pool_sizes[qi] = size;
    }
    if (a.exchange.n_shards > 0) {
      // shard exchange: deliver this list to the query's owner; the last shard to arrive merges (exchange.cuh).
      // The warp's whole shared-memory region is free by now and serves as merge scratch.
      __syncwarp();
      exchange_epilogue(a.exchange, qi, a.topk, min(static_cast<uint32_t>(size), a.topk), a.out_ids + size_t(qi) * a.topk,
                        a.out_dist + size_t(qi) * a.topk, reinterpret_cast<uint2 *>(wbase), lane);
    }
    if (vs.ovf_used) visit_overflow_release(ovf_slot(a), a.ovf_size, vs.listed, vs.big, a.big, a.big_words, a.big_flags);
    if (lane == 0) {
      atomicAdd(&a.stats[0], static_cast<unsigned long long>(ndc));
      atomicAdd(&a.stats[1], static_cast<unsigned long long>(hops));
      if (vs.ovf_used) {
        atomicAdd(&a.stats[3], 1ull);
        atomicAdd(&a.stats[5], static_cast<unsigned long long>(vs.listed));
      }
      if (vs.big >= 0) atomicAdd(&a.stats[4], 1ull);
    }
    __syncwarp();
  }
}

#undef AB_EVAL

template <int QUANT, bool IP, int TPV, int CPL, bool EXACT, int INFL = AB_INFLIGHT>
void launch_one(const SearchPlan &p, const SearchArgs &a, cudaStream_t s) {
  auto kern = beam_search_kernel<QUANT, IP, TPV, CPL, EXACT, INFL>;
  AB_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(p.smem_bytes)));
  // The unified L1 / shared-memory array also stages the row loads in flight: ask for exactly the shared memory
  // the resident CTAs need and leave the rest to L1.
  const size_t need = size_t(p.ctas_per_sm) * (p.smem_bytes + 1024);
  const int pct = static_cast<int>(std::min<size_t>(100, (need * 100 + 228 * 1024 - 1) / (228 * 1024)));
  AB_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributePreferredSharedMemoryCarveout, pct));
  kern<<<dim3(p.ctas), dim3(p.warps_per_cta * 32), p.smem_bytes, s>>>(a);
}


}  // namespace

}  // namespace ab
