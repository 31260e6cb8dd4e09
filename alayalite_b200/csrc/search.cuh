// Host-visible description of one batched graph search launch (search.cu).
#pragma once
#include "common.cuh"
#include "exchange.cuh"

namespace ab {

// Everything the traversal kernel reads, by value. Device pointers unless noted.
// Multi-GPU: see exchange.cuh (the kernel that produces the final lists also delivers them to the owner shard of
// each query and the last shard to arrive merges).
constexpr uint32_t kSpillList = 4096;  // overflow-table entries remembered per query for the targeted clean-up
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
  // Host callers with pageable buffers: the queries are still arriving in chunks of q_chunk rows over a copy stream
  // while this kernel runs; chunk c is complete once q_ready[c] != 0 (written by the copy engine behind the data).
  // null = everything is there.
  const uint32_t *q_ready;
  uint32_t q_chunk;
  uint32_t nq;
  uint32_t ef, topk;
  // outputs: first topk pool entries (ids 0 / dist 0 past the pool size, like the reference's
  // zero-initialised LinearPool), optional whole pools
  uint32_t *out_ids;   // [nq, topk]
  float *out_dist;     // [nq, topk] or null
  uint32_t *pool_ids;  // [nq, ef] or null
  float *pool_dist;    // [nq, ef] or null
  uint32_t *pool_sizes;
  Exchange exchange;   // raw spaces: the traversal kernel delivers the lists (SQ spaces: the re-rank kernel does)
  int pad_empty;       // out_ids slots past the pool size get 0xFFFFFFFF instead of the reference's 0 (input of a
                       // re-rank that feeds an exchange: padding must not masquerade as row 0)
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
  uint32_t hash_buckets;        // > 0: two-choice table of this many buckets of four 16-bit entries (see visit_buckets)
  // Visited ids that do not fit the shared-memory table continue in the warp's private overflow table in global
  // memory: [slots, ovf_size + kSpillList] words per slot -- an open-addressing table of ids (all 0xFFFFFFFF between
  // queries) followed by a scratch list of the entries a query claimed, so that giving the table back clean costs
  // a few hundred word stores. Its size follows ef, not the number of rows. A query that fills even that (3/4
  // load) claims one of n_big pooled N-bit bitmaps (the reference's own structure) for the rest of its run.
  uint32_t *ovf;
  uint32_t ovf_size;            // entries per slot, power of two
  uint32_t ovf_limit;           // ids in the overflow table past which the query moves on to a pooled bitmap
  uint32_t *big;                // [n_big, big_words] bitmaps, all zero while free
  uint32_t *big_flags;          // [n_big] 0 = free, 1 = claimed
  uint32_t n_big;
  size_t big_words;
  // bit 0: rows of the later passes of a hop are requested from the TMA unit (bulk L2 prefetch) while the first pass
  // is evaluated; bit 1: the rows of every neighbour of the node being expanded are requested before the visited
  // filter runs (speculative: rows of already visited neighbours are fetched too)
  int prefetch;
  unsigned long long *stats;    // {ndc level 0, hops, ndc overlay, queries that used the overflow table,
                                //  queries that went on to a pooled bitmap, ids put into overflow tables}
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
  Exchange exchange;         // n_shards > 0: lists delivered to the owner shards
  int no_zero_fill;          // per-shard lists that will be merged: only the topk real candidates are re-ranked (the
                             // reference's (ef - topk) zero slots would put row 0 of every shard into the merge)
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
  uint32_t hash_buckets;  // > 0: two-choice bucket format
  size_t smem_bytes;
  int slots;          // ctas * warps_per_cta (overflow-table slots needed)
  int max_slots;      // slots any plan for this batch size could need (workspace sizing)
  int deep;           // 1: the six-in-flight kernel variant (raw rows of 24 / 32 chunks, at most 16 / 12 resident warps)
  uint32_t ovf_size;  // entries of one overflow table (power of two)
};

// Global-memory part of the visited set shared by every launch of one owner (an index, a graph build): one overflow
// table per resident warp and a small pool of N-bit bitmaps. Memory follows ef and the number of resident warps
// (about 50 KB per warp at ef 64), not the number of rows; only the pooled bitmaps (8 by default) are N bits each.
struct VisitedWorkspace {
  DevBuf<uint32_t> ovf, big, big_flags;
  uint32_t ovf_size = 0;
  int slots = 0;
  uint32_t n_big = 0;
  size_t big_words = 0;
  // (re)allocate and initialise when the plan needs more than what is there; stream-ordered on `s`
  void prepare(const SearchPlan &plan, uint64_t n_rows, cudaStream_t s);
  void fill(SearchArgs &a) const;
  size_t bytes() const { return (ovf.n + big.n + big_flags.n) * sizeof(uint32_t); }
};

// expected visited ids per query divided by ef on SIFT-like graphs
float visited_model(uint32_t ef);
// visit_ratio: visited ids per query / ef measured on earlier searches of this index, as a multiple of
// visited_model(ef) at the ef it was measured with (0 = unknown)
// min_warp_bytes: shared memory every warp must own at least (scratch of the shard exchange)
SearchPlan plan_search(int quant, int nchunks, int table_len, float visit_ratio, uint32_t ef, uint32_t nq, uint32_t n_rows,
                       int sm_count, size_t smem_limit, size_t min_warp_bytes = 0);
void launch_search(int quant, const SearchPlan &plan, const SearchArgs &args, cudaStream_t stream);
// per-quantisation translation units (search_sq8.cu / search_sq4.cu); args already completed by launch_search
void launch_search_sq8(bool ip, const SearchPlan &plan, const SearchArgs &args, cudaStream_t stream);
void launch_search_sq4(bool ip, const SearchPlan &plan, const SearchArgs &args, cudaStream_t stream);
void launch_rerank(const RerankArgs &args, int sm_count, cudaStream_t stream);

}  // namespace ab
