// Batched beam search over the level-0 graph (+ overlay descent) and the SQ re-rank.
//
// Replaces, on the device (paths relative to /root/reference):
//   GraphSearchJob::search / search_solo   include/executor/jobs/graph_search_job.hpp:221-258, :302-335
//   OverlayGraph::initialize               include/index/graph/overlay_graph.hpp:122-144
//   LinearPool (insert / pop / cursor)     include/utils/query_utils.hpp:237-312
//   DynamicBitset visited set              include/utils/query_utils.hpp:69-115
//   Scheduler / Worker coroutine executor  include/executor/scheduler.hpp:113-203, worker.hpp:111-136
//   PyIndex::rerank                        python/include/index.hpp:450-488
//
// This file: the re-rank kernel, the launch plan (occupancy / shared-memory carve-out / visited-table format) and
// the dispatch of the raw-row kernel variants. The traversal kernel itself is in search_kernel.cuh (design notes
// there and in DESIGN.md 4.1); its SQ instantiations are in search_sq8.cu / search_sq4.cu.
#include "search_kernel.cuh"

namespace ab {

namespace {

// PyIndex::rerank (python/include/index.hpp:450-467): exact raw-space distance for the ef entries of
// res_pool -- the topk ids search wrote plus (ef - topk) untouched zeros -- ordered by (dist, id).
__host__ __device__ inline size_t rerank_warp_bytes(const RerankArgs &a) {
  const uint32_t nz = a.no_zero_fill ? 0u : (a.ef - a.topk < a.topk ? a.ef - a.topk : a.topk);
  const size_t own = size_t(a.nchunks) * 16 + round_up16(size_t(a.topk + nz) * 4) * 2;
  const size_t scratch = round_up16(size_t(a.exchange.n_shards) * a.topk * 8);
  return own > scratch ? own : scratch;
}

template <bool IP, int TPV, int MAXC>
__global__ void __launch_bounds__(128) rerank_kernel(const RerankArgs a) {
  extern __shared__ __align__(16) uint8_t smem[];
  const int lane = lane_id();
  const int warp = threadIdx.x >> 5;
  const int warps = blockDim.x >> 5;
  const size_t q_bytes = size_t(a.nchunks) * 16;
  // copies of id 0 that can reach the output (none when the lists feed a shard merge: padding is not a row)
  const uint32_t nz = a.no_zero_fill ? 0u : min(a.ef - a.topk, a.topk);
  const uint32_t L = a.topk + nz;
  const size_t per_warp = rerank_warp_bytes(a);
  uint8_t *wbase = smem + size_t(warp) * per_warp;
  float *q_raw = reinterpret_cast<float *>(wbase);
  float *ld = reinterpret_cast<float *>(wbase + q_bytes);
  uint32_t *li = reinterpret_cast<uint32_t *>(wbase + q_bytes + round_up16(size_t(L) * 4));
  QuerySmem qs;
  qs.q_raw = q_raw;
  qs.q_code = nullptr;
  qs.scale = nullptr;
  qs.vmin = nullptr;
  constexpr int G = 32 / TPV;
  const int g = lane / TPV, lg = lane % TPV;

  for (uint32_t qi = blockIdx.x * warps + warp; qi < a.nq; qi += gridDim.x * warps) {
    const float *qg = a.queries + size_t(qi) * a.dim;
    const int nfl = a.nchunks * 4;
    float ss = 0.f;
    for (int i = lane; i < nfl; i += 32) {
      const float x = i < static_cast<int>(a.dim) ? qg[i] : 0.f;
      q_raw[i] = x;
      ss = fmaf(x, x, ss);
    }
    if (a.metric == AB_METRIC_COS) {
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) ss += __shfl_xor_sync(kFull, ss, o);
      const float inv = static_cast<float>(1.0 / static_cast<double>(sqrtf(ss)));
      __syncwarp();
      for (int i = lane; i < nfl; i += 32) q_raw[i] *= inv;
    }
    for (uint32_t i = lane; i < L; i += 32) li[i] = i < a.topk ? a.cand_ids[size_t(qi) * a.topk + i] : 0u;
    __syncwarp();
    // distances of the topk candidates and of id 0 (entry topk), shared by all its copies
    const uint32_t ndist = a.topk + (nz > 0 ? 1u : 0u);
    for (uint32_t base = 0; base < ndist; base += G) {
      const uint32_t k = base + g;
      const bool act = k < ndist;
      const uint32_t cv = act ? li[k] : 0u;
      const uint32_t cr = cv < a.n_rows ? cv : 0u;  // ids the raw space does not hold score FLT_MAX below; never dereferenced
      float d = row_distance<AB_QUANT_NONE, IP, TPV, MAXC>(reinterpret_cast<const uint8_t *>(a.raw) + size_t(cr) * a.raw_stride,
                                                          a.nchunks, lg, qs, act);
      if (act && lg == 0) {
        // RawSpace::QueryComputer checks the raw storage's validity bit (raw_space.hpp:297-300): rows the
        // raw space never received (SQ inserts, graph_update_job.hpp:75) and tombstones score FLT_MAX
        if (cv >= a.n_rows || (a.valid != nullptr && ((a.valid[cv >> 5] >> (cv & 31)) & 1u) == 0)) d = FLT_MAX;
        ld[k] = d;
      }
    }
    __syncwarp();
    for (uint32_t i = a.topk + 1 + lane; i < L; i += 32) ld[i] = ld[a.topk];
    __syncwarp();
    // rank sort by (dist, id, position)
    for (uint32_t i = lane; i < L; i += 32) {
      const float di = ld[i];
      const uint32_t ii = li[i];
      uint32_t rank = 0;
      for (uint32_t j = 0; j < L; j++) {
        const float dj = ld[j];
        const uint32_t ij = li[j];
        rank += (dj < di || (dj == di && (ij < ii || (ij == ii && j < i)))) ? 1u : 0u;
      }
      if (rank < a.topk) {
        a.out_ids[size_t(qi) * a.topk + rank] = ii;
        if (a.out_dist) a.out_dist[size_t(qi) * a.topk + rank] = di;
      }
    }
    __syncwarp();
    if (a.exchange.n_shards > 0) {
      // candidates the traversal padded with 0xFFFFFFFF (pool smaller than topk) sort last: the real ones are a prefix
      uint32_t nvalid = 0;
      for (uint32_t i = lane; i < a.topk; i += 32) nvalid += li[i] != kEmptyId ? 1u : 0u;
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) nvalid += __shfl_xor_sync(kFull, nvalid, o);
      __syncwarp();
      exchange_epilogue(a.exchange, qi, a.topk, nvalid, a.out_ids + size_t(qi) * a.topk, a.out_dist + size_t(qi) * a.topk,
                        reinterpret_cast<uint2 *>(wbase), lane);
    }
  }
}

int pick_tpv(int nchunks) {
  // aim for up to 8 x 128-bit loads in flight per lane
  if (nchunks > 128) return 32;
  if (nchunks > 64) return 16;
  if (nchunks > 32) return 8;
  return 4;
}

// raw space: register-resident query, TPV * CPL chunks per row
template <bool IP>
bool launch_raw_fast(const SearchPlan &p, const SearchArgs &a, cudaStream_t s) {
  const int n = a.nchunks;
#define AB_FAST(T, C)                                                   \
  if (n == T * C) { launch_one<AB_QUANT_NONE, IP, T, C, true>(p, a, s); return true; }
#define AB_FAST_LE(T, C)                                                \
  if (n <= T * C) { launch_one<AB_QUANT_NONE, IP, T, C, false>(p, a, s); return true; }
  if (p.deep && n == 24) { launch_one<AB_QUANT_NONE, IP, 8, 3, true, 6>(p, a, s); return true; }
  if (p.deep && n == 32) { launch_one<AB_QUANT_NONE, IP, 8, 4, true, 6>(p, a, s); return true; }
  AB_FAST(8, 3)    // d = 96
  AB_FAST(8, 4)    // d = 128
  AB_FAST(16, 4)   // d = 256
  AB_FAST(32, 4)   // d = 512
  AB_FAST(32, 6)   // d = 768
  AB_FAST_LE(8, 2)
  AB_FAST_LE(8, 4)
  AB_FAST_LE(16, 4)
  AB_FAST_LE(32, 4)
  AB_FAST_LE(32, 8)  // up to d = 1024 (d = 960 lands here)
#undef AB_FAST
#undef AB_FAST_LE
  return false;
}

// raw rows: the three-in-flight / register-query variants above, or the generic kernel for rows above 256 chunks
template <bool IP>
void launch_search_raw(const SearchPlan &p, const SearchArgs &a, cudaStream_t s) {
  if (!launch_raw_fast<IP>(p, a, s)) launch_one<AB_QUANT_NONE, IP, 32, 0, false>(p, a, s);
  AB_CUDA(cudaGetLastError());
  count_launch();
}

}  // namespace

void VisitedWorkspace::prepare(const SearchPlan &plan, uint64_t n_rows, cudaStream_t s) {
  static const uint32_t big_slots = getenv("AB_BIG_SLOTS") ? static_cast<uint32_t>(std::max(1, atoi(getenv("AB_BIG_SLOTS")))) : 8u;
  if (plan.ovf_size > ovf_size || plan.slots > slots) {
    // sized for the most warps any plan can keep resident (24 per SM), so that a plan change between launches --
    // the visited-ids feedback, another ef -- does not reallocate inside a caller's timed loop
    ovf_size = std::max(ovf_size, plan.ovf_size);
    slots = std::max(std::max(slots, plan.slots), plan.max_slots);
    const size_t words = (size_t(ovf_size) + kSpillList) * size_t(slots);
    ovf.reserve(words);
    AB_CUDA(cudaMemsetAsync(ovf.p, 0xFF, words * sizeof(uint32_t), s));
  }
  const size_t words = round_up((n_rows + 31) / 32, 4);
  if (words > big_words || n_big != big_slots) {
    big_words = std::max(big_words, words);
    n_big = big_slots;
    big.alloc_zero(big_words * n_big, s);
    big_flags.alloc_zero(n_big, s);
  }
}

void VisitedWorkspace::fill(SearchArgs &a) const {
  a.ovf = ovf.p;
  a.ovf_size = ovf_size;
  a.ovf_limit = ovf_size - ovf_size / 4;
  a.big = big.p;
  a.big_flags = big_flags.p;
  a.n_big = n_big;
  a.big_words = big_words;
}

float visited_model(uint32_t ef) { return 16.0f + 400.0f / static_cast<float>(std::max(1u, ef)); }

static SearchPlan plan_search_impl(int quant, int nchunks, int table_len, float visit_ratio, uint32_t ef, uint32_t nq, uint32_t n_rows,
                                   int sm_count, size_t smem_limit, size_t min_warp_bytes, int reg_cap, int wpc_cap);

SearchPlan plan_search(int quant, int nchunks, int table_len, float visit_ratio, uint32_t ef, uint32_t nq, uint32_t n_rows,
                       int sm_count, size_t smem_limit, size_t min_warp_bytes) {
  SearchPlan p = plan_search_impl(quant, nchunks, table_len, visit_ratio, ef, nq, n_rows, sm_count, smem_limit, min_warp_bytes, 64, 8);
  // Deep variant (six candidate rows per lane group in flight instead of three): when shared memory -- the visited
  // table of a large ef -- leaves room for no more warps than the deep kernel's registers allow anyway, the same
  // warps carry twice the rows in flight. d = 96: 16 warps (128 registers), d = 128: 12 warps of 6 per CTA (168).
  static const int deep_on = getenv("AB_DEEP") ? atoi(getenv("AB_DEEP")) : 1;
  if (deep_on && quant == AB_QUANT_NONE && (nchunks == 24 || nchunks == 32)) {
    const int cap = nchunks == 24 ? 16 : 12, wpc = nchunks == 24 ? 8 : 6;
    const int have = p.warps_per_cta * p.ctas_per_sm;
    // small batches (every query has a warp of its own anyway): latency per hop is all that matters
    const bool small_batch = nq <= static_cast<uint32_t>(sm_count) * static_cast<uint32_t>(cap);
    if (deep_on == 2 || have <= cap || small_batch) {
      SearchPlan d = plan_search_impl(quant, nchunks, table_len, visit_ratio, ef, nq, n_rows, sm_count, smem_limit, min_warp_bytes, cap, wpc);
      if (deep_on == 2 || small_batch || 2 * d.warps_per_cta * d.ctas_per_sm > have + have / 4) {
        d.deep = 1;
        return d;
      }
    }
  }
  return p;
}

static SearchPlan plan_search_impl(int quant, int nchunks, int table_len, float visit_ratio, uint32_t ef, uint32_t nq, uint32_t n_rows,
                                   int sm_count, size_t smem_limit, size_t min_warp_bytes, int reg_cap, int wpc_cap) {
  SearchPlan p{};
  p.tpv = pick_tpv(nchunks);
  static const float hash_factor = getenv("AB_HASH_FACTOR") ? static_cast<float>(atof(getenv("AB_HASH_FACTOR"))) : 1.7f;
  static const int max_warps = getenv("AB_MAX_WARPS") ? atoi(getenv("AB_MAX_WARPS")) : 64;
  static const int hash16 = getenv("AB_HASH16") ? atoi(getenv("AB_HASH16")) : 1;
  // Hash entries per warp so that nearly every query stays below the 3/4 spill mark. Visited ids per query grow
  // like ratio(ef) * ef with ratio(ef) ~ 16 + 400 / ef on SIFT-like graphs (31 at ef 20, 23.6 at ef 52, 20.9 at ef 100,
  // 18.7 at ef 200, measured); visit_ratio is this index's own measured multiple of that model (0 = unknown, 1.0
  // assumed; harder data has more). 1.7 = 1 / 0.75 load x 1.28 for the spread between queries.
  const float ratio = visited_model(ef) * (visit_ratio > 0.f ? std::min(4.0f, std::max(0.75f, visit_ratio)) : 1.0f);
  uint32_t h = static_cast<uint32_t>(ratio * hash_factor * static_cast<float>(ef));
  h = std::min(32768u, std::max(512u, (h + 63u) & ~63u));
  const size_t cta_fixed = quant != AB_QUANT_NONE ? size_t(table_len) * 8 : 0;
  const size_t pool_bytes = round_up16(size_t(ef + 1) * 8);
  const size_t other = 256 + size_t(nchunks) * 16 + pool_bytes;

  // 16-bit entries if the id range is small enough for (displacement, remainder) to fit (see visit())
  int idbits = 1;
  while (idbits < 31 && (1ull << idbits) < n_rows) idbits++;
  auto quotient = [&](uint32_t entries, SearchPlan &q) {
    q.hash_rbits = 0;
    if (!hash16 || entries >= (1ull << idbits)) return false;
    const uint64_t S = uint64_t(entries) << (32 - idbits);  // scaled scramble units per table slot
    const uint32_t M = static_cast<uint32_t>((1ull << 32) / S) + 1;
    int rbits = 1;
    while ((1u << rbits) < M) rbits++;
    // Down to two displacement bits the quotient entries still win: a probe that runs out of displacement range
    // continues in the warp's bitmap, which is cheap to give back since queries list the ids they put there
    // (12.5 M rows, ef 64: 1.93 ms with 32-bit ids, 1.41 ms with 14 + 2 bit entries).
    static const int rbits_max = getenv("AB_HASH_RBITS_MAX") ? atoi(getenv("AB_HASH_RBITS_MAX")) : 14;
    static const int dbits_cap = getenv("AB_HASH_DBITS_CAP") ? atoi(getenv("AB_HASH_DBITS_CAP")) : 16;  // test knob
    if (rbits > rbits_max) return false;  // too few displacement bits left: keep 32-bit entries
    q.hash_rbits = rbits;
    q.hash_xshift = 32 - idbits;
    q.hash_rmul = M;
    q.hash_dlimit = ((1u << std::min(16 - rbits, std::max(1, dbits_cap))) - 1) << rbits;
    return true;
  };
  // Two-choice bucket format (visit_buckets): buckets of four 16-bit entries, an entry = which of the id's two
  // buckets it sits in (1 bit) + the remainder of that scramble within the bucket (15 bits, 0xFFFF = empty). Needs
  // at most 32767 ids of the id space per bucket; otherwise the slot format below, otherwise 32-bit ids.
  static const int buckets_on = getenv("AB_HASH_BUCKETS") ? atoi(getenv("AB_HASH_BUCKETS")) : 1;
  static const int buckets_minv = getenv("AB_HASH_BUCKETS_MINV") ? atoi(getenv("AB_HASH_BUCKETS_MINV")) : 0;
  auto quotient_slots = quotient;
  auto quotient_any = [&](uint32_t entries, SearchPlan &q) {
    q.hash_buckets = 0;
    if (buckets_on && ratio * static_cast<float>(ef) >= static_cast<float>(buckets_minv) && hash16 && entries >= 64 && entries < (1ull << idbits)) {
      const uint32_t nb = entries / 4;
      const uint64_t S = uint64_t(nb) << (32 - idbits);
      const uint64_t M = (1ull << 32) / S + 1;
      if (M <= 32767) {
        q.hash_buckets = nb;
        q.hash_rbits = 15;
        q.hash_xshift = 32 - idbits;
        q.hash_rmul = static_cast<uint32_t>(M);
        q.hash_dlimit = 0;
        return true;
      }
    }
    return quotient_slots(entries, q);
  };
  uint32_t entry_bytes = 4;
  if (quotient_any(std::max(h, 2048u), p)) {  // 16-bit entries are cheap: never below 2048 of them (4 KB)
    h = std::max(h, 2048u);
    entry_bytes = 2;
  }
  // resident warps the register file allows: the three-in-flight raw kernels are built for 3 x 256 threads,
  // everything else (SQ spaces, rows above 128 chunks) for 2 x 256
  const int reg_warps = std::min(std::min(max_warps, reg_cap), (quant == AB_QUANT_NONE && nchunks <= 128) ? 8 * AB_MINBLOCKS : 16);
  // Full register-limited occupancy under a 196 KB carve-out is worth a tighter table (the step from 24 to
  // 20 warps costs ~25 % at ef 100): shrink down to 1.3 / 1.7 of the wanted size for it. Below 12 resident
  // warps, shrink further and let long queries continue in their global bitmap.
  auto fit_entries = [&](size_t limit, int warps) -> uint32_t {
    const size_t ctas = (warps + 7) / 8;
    const size_t total = ctas * (cta_fixed + 1024) + size_t(warps) * other;
    if (limit <= total) return 0;
    return static_cast<uint32_t>((limit - total) / warps / entry_bytes) & ~63u;
  };
  // The raw kernels leave ~60 KB of the unified array to L1 (section "Shared memory vs. L1" of DESIGN.md). The SQ
  // kernels are capped at 16 warps by their registers anyway and gain more from a larger visited table than from L1
  // (SQ8 d = 960, ef 180: 13.9 -> 12.7 ms with the whole carve-out): they may use all of it.
  static const size_t soft_kb = getenv("AB_SMEM_SOFT_KB") ? static_cast<size_t>(atoi(getenv("AB_SMEM_SOFT_KB"))) : 0;
  const size_t soft_limit = std::min(smem_limit, (soft_kb ? soft_kb : (quant != AB_QUANT_NONE ? size_t(227) : size_t(196))) * 1024);
  const uint32_t h_full = fit_entries(soft_limit, reg_warps);
  const uint32_t h_min = std::max(entry_bytes == 2 ? 2048u : 512u, (static_cast<uint32_t>(float(h) * (1.3f / 1.7f)) + 63u) & ~63u);
  if (h_full < h && h_full >= h_min) {
    h = h_full;
  } else {
    const uint32_t h12 = fit_entries(smem_limit, 12);
    if (h12 < h) h = std::max(512u, h12);
  }
  entry_bytes = quotient_any(h, p) ? 2 : 4;
  // the shard exchange merges in the warp's own shared memory: make sure the region is large enough
  if (other + size_t(h) * entry_bytes < min_warp_bytes) {
    h = static_cast<uint32_t>((min_warp_bytes - other + entry_bytes - 1) / entry_bytes + 63) & ~63u;
    entry_bytes = quotient_any(h, p) ? 2 : 4;
  }
  p.hash_size = h;
  p.hash_words = h * entry_bytes / 4;
  const size_t pw = other + size_t(p.hash_words) * 4;
  // CTA size: whatever packs the most warps into the SM (each CTA also pays cta_fixed + 1 KB of reserved
  // shared memory). The unified L1 / shared-memory array also stages the row loads in flight, and the step
  // from a 196 KB to a 228 KB carve-out halves what is left for that (60 -> 28 KB): stay at or below 196 KB
  // unless that costs more than a fifth of the resident warps.
  auto pack = [&](size_t limit, int &wpc_out, int &ctas_out) {
    wpc_out = 1;
    ctas_out = 0;
    for (int wpc = wpc_cap; wpc >= 1; --wpc) {
      const int by_smem = static_cast<int>(limit / (cta_fixed + size_t(wpc) * pw + 1024));
      const int ctas = std::min(by_smem, reg_warps / wpc);
      if (ctas * wpc > ctas_out * wpc_out) {
        wpc_out = wpc;
        ctas_out = ctas;
      }
    }
  };
  int wpc_a, ctas_a, wpc_b, ctas_b;
  pack(soft_limit, wpc_a, ctas_a);
  pack(smem_limit, wpc_b, ctas_b);
  const bool small = ctas_a > 0 && 5 * wpc_a * ctas_a >= 4 * wpc_b * ctas_b;
  const int best_wpc = small ? wpc_a : wpc_b, best_ctas = std::max(1, small ? ctas_a : ctas_b);
  p.warps_per_cta = best_wpc;
  p.smem_bytes = cta_fixed + size_t(best_wpc) * pw;
  p.ctas_per_sm = best_ctas;
  const int want_ctas = static_cast<int>((nq + best_wpc - 1) / best_wpc);
  p.ctas = std::max(1, std::min(want_ctas, sm_count * best_ctas));
  p.slots = p.ctas * p.warps_per_cta;
  p.max_slots = nq >= static_cast<uint32_t>(sm_count) * 24u ? sm_count * 24 : p.slots;
  // overflow table: room for three times the ids a query is expected to visit, so that the pooled bitmaps stay a
  // rarely used safety net even when the shared-memory table takes next to nothing
  static const uint32_t ovf_forced = getenv("AB_OVF_SIZE") ? static_cast<uint32_t>(atoi(getenv("AB_OVF_SIZE"))) : 0u;
  // (from the model alone, not from this index's measured ratio: the feedback must not reallocate the workspace
  // in the middle of a caller's loop)
  uint32_t want = ovf_forced ? ovf_forced : static_cast<uint32_t>(std::min(5.0f * visited_model(ef) * static_cast<float>(ef), 4.0e6f));
  uint32_t o = ovf_forced ? 64u : 4096u;
  while (o < want) o <<= 1;
  p.ovf_size = o;
  // (the sticky all-warps sizing is for the common small tables only: at a very large ef few warps are resident and
  // their tables are megabytes each)
  if (o > 16384) p.max_slots = p.slots;
  return p;
}

void launch_search(int quant, const SearchPlan &plan, const SearchArgs &args_in, cudaStream_t stream) {
  SearchArgs args = args_in;
  args.hash_size = plan.hash_size;
  args.hash_words = plan.hash_words;
  static const int load_pct = getenv("AB_HASH_LOAD") ? std::min(95, std::max(10, atoi(getenv("AB_HASH_LOAD")))) : 75;
  args.hash_limit = static_cast<uint32_t>(uint64_t(plan.hash_size) * load_pct / 100);  // freeze the table at this load
  args.hash_buckets = plan.hash_buckets;
  if (plan.hash_buckets) args.hash_limit = 0xFFFFFFFFu;  // the bucket format needs no freeze: an id whose two buckets are full moves on
  args.hash_rbits = plan.hash_rbits;
  args.hash_xshift = plan.hash_xshift;
  args.hash_rmul = plan.hash_rmul;
  args.hash_dlimit = plan.hash_dlimit;
  // Long code rows (TPV >= 16: six rows per pass, a hop of ~23 new neighbours is four passes): the rows of the later
  // passes are requested from the TMA unit when the first pass starts (SQ8 d = 960, ef 180: 12.1 -> 11.6 ms). Measured
  // and left off elsewhere: SQ4 d = 768 (-1 %), raw d = 128 (+-1 %); the speculative variant (bit 1: every neighbour
  // before the visited filter) costs 1-10 % everywhere.
  const int prefetch = getenv("AB_PREFETCH") ? atoi(getenv("AB_PREFETCH")) : ((quant != AB_QUANT_NONE && plan.tpv >= 16) ? 1 : 0);
  args.prefetch = prefetch;
  const bool ip = args.metric != AB_METRIC_L2;
  if (quant == AB_QUANT_NONE) {
    ip ? launch_search_raw<true>(plan, args, stream) : launch_search_raw<false>(plan, args, stream);
  } else if (quant == AB_QUANT_SQ8) {
    launch_search_sq8(ip, plan, args, stream);
  } else {
    launch_search_sq4(ip, plan, args, stream);
  }
}

void launch_rerank(const RerankArgs &a, int sm_count, cudaStream_t stream) {
  const int tpv = pick_tpv(a.nchunks);
  const int warps = 4;
  const size_t per_warp = rerank_warp_bytes(a);
  const size_t smem = per_warp * warps;
  const int ctas = std::max(1, std::min<int>((a.nq + warps - 1) / warps, sm_count * 8));
  const bool ip = a.metric != AB_METRIC_L2;
#define AB_RR(IPV, TPV)                                                                                   \
  {                                                                                                       \
    auto kern = rerank_kernel<IPV, TPV, 8>;                                                               \
    AB_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem))); \
    kern<<<ctas, warps * 32, smem, stream>>>(a);                                                          \
  }
  if (ip) {
    switch (tpv) {
      case 4: AB_RR(true, 4); break;
      case 8: AB_RR(true, 8); break;
      case 16: AB_RR(true, 16); break;
      default: AB_RR(true, 32); break;
    }
  } else {
    switch (tpv) {
      case 4: AB_RR(false, 4); break;
      case 8: AB_RR(false, 8); break;
      case 16: AB_RR(false, 16); break;
      default: AB_RR(false, 32); break;
    }
  }
#undef AB_RR
  AB_CUDA(cudaGetLastError());
  count_launch();
}

}  // namespace ab
