// Host-visible description of one batched graph search launch (search.cu).
#pragma once
#include "common.cuh"

namespace ab {

// Everything the traversal kernel reads, by value. Device pointers unless noted.
// Multi-GPU exchange fused into the kernel that produces the final lists: besides out_ids / out_dist the
// results (ids + id_offset) are stored straight into the [slots, nq, topk] gather buffers of every rank of
// the box -- peer memory mapped over NVLink -- at slot `slot`, so the transfer overlaps the search.
constexpr int kMaxDest = 16;
constexpr uint32_t kSpillList = 4096;  // ids remembered per query for the targeted bitmap clean-up
struct Scatter {
  int n_dest;  // 0 = off
  uint32_t slot;
  uint32_t id_offset;
  uint32_t *ids[kMaxDest];
  float *dist[kMaxDest];
};

struct SearchArgs {
  // search space rows: raw floats or SQ codes, row stride in bytes (multiple of 64)
  const uint8_t *rows;
  uint32_t row_stride;
  int nchunks;         // 16-byte chunks per row that carry data
  int table_len;       // entries in the dequantisation tables (SQ only)
  const float *qmin;   // [dim] quantizer min (SQ only)
  const float *qmax;   // [dim] quantizer max (SQ only)
  const float *scale;  // [table_len] (max-min)/levels, zero padded (SQ only)
  const float *vmin;   // [table_len] min, zero padded (SQ only)
  // graph
  const uint32_t *adj;       // [n, 32] level-0 adjacency, 0xFFFFFFFF padded, compact from the left
  const uint32_t *ov_off;    // [n] first overlay row of a node (rows of 32 ids), 0xFFFFFFFF = level 0; null = no overlay
  const uint32_t *ov_lists;  // overlay rows
  const uint32_t *valid;     // raw-space validity bitmap (null = every row valid)
  uint32_t ep, ep_level, n, dim;
  int metric;  // ab_metric
  // queries
  const float *queries;   // query i = queries + (q_ids ? q_ids[i] : i) * q_stride
  const uint32_t *q_ids;  // optional indirection (graph build: the queries are base rows)
  size_t q_stride;        // floats between consecutive query rows
  int q_prenormalized;    // COS only: queries are stored rows, already normalised
  uint32_t nq;
  uint32_t ef, topk;
  // outputs: first topk pool entries (ids 0 / dist 0 past the pool size, like the reference's
  // zero-initialised LinearPool), optional whole pools
  uint32_t *out_ids;   // [nq, topk]
  float *out_dist;     // [nq, topk] or null
  uint32_t *pool_ids;  // [nq, ef] or null
  float *pool_dist;    // [nq, ef] or null
  uint32_t *pool_sizes;
  Scatter scatter;
  // scheduling + visited set
  uint32_t *work_counter;       // zeroed before launch
  // visited-set table, filled in by launch_search from the plan
  uint32_t hash_size;           // entries per warp (multiple of 64)
  uint32_t hash_words;          // 32-bit words the table occupies in shared memory
  uint32_t hash_limit;          // freeze the table and continue in the bitmap past this many ids (3/4 load)
  int hash_rbits;               // 0: 32-bit entries holding ids; else 16-bit (displacement, remainder) entries
  int hash_xshift;              // 32 - bits of the id space
  uint32_t hash_rmul;           // remainder = umulhi(low word of scramble * entries, hash_rmul)
  uint32_t hash_dlimit;         // first invalid entry code (displacement range exhausted)
  // per-warp visited bitmaps: [slots, bitmap_words + kSpillList] words. The first bitmap_words of a slot are the
  // bitmap (all zero between launches); the tail is a scratch list of the ids a query put there, so that giving
  // the bitmap back clean costs a few hundred word stores instead of rewriting N / 8 bytes
  uint32_t *bitmaps;
  size_t bitmap_words;
  unsigned long long *stats;    // {ndc level 0, hops, ndc overlay, spilled queries}
};

struct RerankArgs {
  const float *raw;  // raw rows
  size_t raw_stride; // bytes
  int nchunks;
  const uint32_t *valid;
  uint32_t n_rows;  // rows the raw space holds; ids past it score FLT_MAX (not valid in the raw storage)
  const float *queries;
  uint32_t nq, dim, ef, topk;
  int metric;
  const uint32_t *cand_ids;  // [nq, topk] search-space top-k
  uint32_t *out_ids;         // [nq, topk]
  float *out_dist;           // [nq, topk] or null
  Scatter scatter;
};

// Launch configuration chosen on the host (search.cu)
struct SearchPlan {
  int tpv;            // lanes per candidate row
  int warps_per_cta;
  int ctas;
  int ctas_per_sm;    // resident CTAs per SM the plan counts on
  uint32_t hash_size;   // visited-table entries per warp
  uint32_t hash_words;  // 32-bit words of shared memory the table takes
  int hash_rbits, hash_xshift;
  uint32_t hash_rmul, hash_dlimit;
  size_t smem_bytes;
  int slots;          // ctas * warps_per_cta (bitmap slots needed)
};

// expected visited ids per query divided by ef on SIFT-like graphs
float visited_model(uint32_t ef);
// visit_ratio: visited ids per query / ef measured on earlier searches of this index, as a multiple of
// visited_model(ef) at the ef it was measured with (0 = unknown)
SearchPlan plan_search(int quant, int nchunks, int table_len, float visit_ratio, uint32_t ef, uint32_t nq, uint32_t n_rows,
                       int sm_count, size_t smem_limit);
void launch_search(int quant, const SearchPlan &plan, const SearchArgs &args, cudaStream_t stream);
// per-quantisation translation units (search_sq8.cu / search_sq4.cu); args already completed by launch_search
void launch_search_sq8(bool ip, const SearchPlan &plan, const SearchArgs &args, cudaStream_t stream);
void launch_search_sq4(bool ip, const SearchPlan &plan, const SearchArgs &args, cudaStream_t stream);
void launch_rerank(const RerankArgs &args, int sm_count, cudaStream_t stream);

}  // namespace ab
