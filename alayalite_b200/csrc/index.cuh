// Device-resident index object behind the C ABI (index.cu, io.cu, build.cu, update.cu).
#pragma once
#include <mutex>
#include <unordered_map>
#include <unordered_set>
#include <vector>

#include "common.cuh"
#include "search.cuh"

struct ab_index {
  ab_params params{};
  int device = 0;
  int sm_count = 148;
  size_t smem_limit = 227 * 1024;
  cudaStream_t stream = nullptr;
  cudaStream_t copy_stream = nullptr;  // host callers with pageable query buffers: chunked copy-in overlapped with the search
  cudaEvent_t copy_event = nullptr;
  uint32_t *h_one = nullptr;           // pinned constant 1 (source of the per-chunk "ready" flags)
  std::mutex mu;

  // ---- spaces (device) -----------------------------------------------------------------------
  uint32_t dim = 0;
  uint64_t n = 0;         // rows in the search space (grows with insert)
  uint64_t n_raw = 0;     // rows in the raw space (== n unless quantised + inserted, reference quirk)
  uint64_t capacity = 0;
  uint32_t raw_stride = 0;   // floats per raw row (64-byte multiple)
  uint32_t code_stride = 0;  // bytes per code row (64-byte multiple)
  uint32_t row_bytes = 0;    // code bytes that carry data
  int raw_chunks = 0, code_chunks = 0, table_len = 0;
  ab::DevBuf<float> raw;          // [capacity, raw_stride]
  ab::DevBuf<uint8_t> codes;      // [capacity, code_stride]
  ab::DevBuf<float> qmin, qmax;   // [dim]
  ab::DevBuf<float> scale, vmin;  // [table_len]
  ab::DevBuf<uint32_t> valid;     // raw-space validity bitmap words (allocated on first remove)
  bool has_removed = false;
  uint32_t delete_cnt_raw = 0, delete_cnt_space = 0;  // RawSpace / SQ space delete_cnt_ (raw_space.hpp:69)

  // ---- graph (device) --------------------------------------------------------------------------
  ab::DevBuf<uint32_t> adj;       // [capacity, 32]
  ab::DevBuf<uint32_t> ov_off;    // [capacity]
  ab::DevBuf<uint32_t> ov_lists;  // [ov_rows, 32]
  uint64_t ov_rows = 0;
  uint32_t ep = 0, ep_level = 0;
  bool has_graph = false, has_overlay = false;
  std::vector<uint32_t> h_levels;  // host copy of the overlay levels (save / get_graph)

  // ---- update context (reference: include/executor/jobs/job_context.hpp:25-29) -----------------
  std::unordered_map<uint32_t, std::vector<uint32_t>> inserted_edges;
  std::unordered_set<uint32_t> removed_vertices;
  std::unordered_map<uint32_t, std::vector<uint32_t>> removed_node_nbrs;
  std::vector<uint8_t> h_valid_raw;    // validity of the raw space rows (host mirror, for save)
  std::vector<uint8_t> h_valid_space;  // validity of the search space rows (SQ) / same as raw
  std::vector<uint8_t> h_valid_graph;  // validity of graph rows

  // ---- search workspace ------------------------------------------------------------------------
  ab::DevBuf<float> w_queries;
  ab::DevBuf<uint32_t> w_ids, w_cand;
  ab::DevBuf<float> w_dist;
  ab::DevBuf<uint32_t> w_pool_ids, w_pool_sizes;
  ab::DevBuf<float> w_pool_dist;
  ab::DevBuf<uint32_t> w_counter;            // [1]
  ab::DevBuf<uint32_t> w_qready;             // [kQueryChunks] per-chunk arrival flags of a pipelined copy-in
  ab::DevBuf<unsigned long long> w_stats;    // [8] device counters of the last launch (SearchArgs::stats)
  ab::VisitedWorkspace w_visited;            // overflow tables + pooled bitmaps of the visited set
  // {ndc, hops, overlay ndc, overflow-table queries, queries, bitmap queries, ids put into overflow tables}
  uint64_t last_stats[7] = {0, 0, 0, 0, 0, 0, 0};
  uint32_t last_ef = 0;      // ef of the launch the device counters belong to
  uint64_t last_nq = 0;      // and its query count
  // lag-free feedback for device-resident callers (no synchronisation on their stream): the counters of launch k
  // are copied to pinned host memory at the start of launch k+1 and picked up once that copy's event has fired
  unsigned long long *h_stats = nullptr;  // pinned [4]
  cudaEvent_t stats_event = nullptr;
  bool stats_pending = false;
  uint32_t stats_ef = 0;
  uint64_t stats_nq = 0;
  float visit_ratio = 0.f;   // visited ids per query / ef over visited_model(ef), from the last counters read back
  // Ordering contract of the shared workspace (work counter, overflow tables, counters, w_cand): every launch that
  // uses it records ws_event behind itself, and the next user -- on whatever stream -- waits for that event first.
  // Searches on one handle therefore never overlap on the device, and insert / remove / build / load order
  // themselves behind every search already enqueued.
  cudaEvent_t ws_event = nullptr;
  bool ws_event_recorded = false;
};

namespace ab {

void index_configure_layout(ab_index *ix, uint32_t dim);
void index_alloc_spaces(ab_index *ix);
void index_upload_graph(ab_index *ix, const uint32_t *adj, const uint32_t *levels, const uint32_t *lists_flat,
                        uint64_t lists_len, uint32_t ep);
void index_download_graph(ab_index *ix, uint32_t *adj, uint32_t *levels, uint32_t *lists_flat);
uint64_t index_lists_len(const ab_index *ix);
// d_queries / outputs are device pointers. cand/pool outputs optional.
void index_search_device(ab_index *ix, const float *d_queries, uint64_t nq, uint32_t topk, uint32_t ef, bool rerank,
                         uint32_t *d_out_ids, float *d_out_dist, uint32_t *d_pool_ids, float *d_pool_dist,
                         uint32_t *d_pool_sizes, cudaStream_t stream, const Exchange *exchange = nullptr, bool pad_empty = false,
                         const uint32_t *q_ready = nullptr, uint32_t q_chunk = 0);
void index_save(ab_index *ix, const char *index_path, const char *data_path, const char *quant_path);
void index_load(ab_index *ix, const char *index_path, const char *data_path, const char *quant_path);
void index_build_graph(ab_index *ix, uint32_t ef_construction);
// NN-Descent k-NN graph + NSG (build.cu)
void index_build_graph_nsg(ab_index *ix, uint32_t ef_construction, uint32_t nnd_iters);
void index_knng_nndescent(ab_index *ix, uint32_t K, uint32_t iters, uint32_t *out_ids);
uint32_t index_insert(ab_index *ix, const float *vec, uint32_t ef);
void index_remove(ab_index *ix, uint32_t id);
void index_sync_valid_bitmap(ab_index *ix);
// make `stream` wait for every search already enqueued on this handle (any stream); see ab_index::ws_event
void index_order_after_searches(ab_index *ix, cudaStream_t stream);
// record the end of a workspace user on `stream`
void index_mark_workspace_use(ab_index *ix, cudaStream_t stream);

}  // namespace ab
