// C ABI of libalaya_b200.so (include/alaya_b200.h): index life cycle, fit, search entry points.
// Host glue only -- the work happens in search.cu, quant.cu, bruteforce.cu, build.cu.
// Mirrors the control flow of PyIndex (/root/reference/python/include/index.hpp:85-506).
#include <algorithm>
#include <cstring>
#include <functional>
#include <memory>

#include "index.cuh"
#include "kernels.cuh"

namespace ab {

std::atomic<uint64_t> g_kernel_launches{0};
static thread_local std::string g_last_error;

template <typename F>
static int guarded(F &&f) {
  try {
    f();
    return AB_OK;
  } catch (const Error &e) {
    g_last_error = e.what();
    return e.code;
  } catch (const std::exception &e) {
    g_last_error = e.what();
    return AB_ERR_INVALID;
  }
}

// for the other translation units of the C ABI (group.cu)
int guarded_call(const std::function<void()> &f) { return guarded(f); }

static void use_device(const ab_index *ix) { AB_CUDA(cudaSetDevice(ix->device)); }

void index_configure_layout(ab_index *ix, uint32_t dim) {
  ix->dim = dim;
  ix->raw_stride = static_cast<uint32_t>(round_up(size_t(dim) * 4, kRowAlign) / 4);
  ix->raw_chunks = static_cast<int>((size_t(dim) * 4 + 15) / 16);
  const int q = ix->params.quantization_type;
  if (q == AB_QUANT_SQ8) {
    ix->row_bytes = dim;
    ix->code_chunks = static_cast<int>((dim + 15) / 16);
    ix->table_len = ix->code_chunks * 16;
  } else if (q == AB_QUANT_SQ4) {
    ix->row_bytes = (dim + 1) / 2;
    ix->code_chunks = static_cast<int>((ix->row_bytes + 15) / 16);
    ix->table_len = ix->code_chunks * 32;
  } else {
    ix->row_bytes = 0;
    ix->code_chunks = 0;
    ix->table_len = 0;
  }
  ix->code_stride = static_cast<uint32_t>(round_up(ix->row_bytes, kRowAlign));
}

void index_alloc_spaces(ab_index *ix) {
  const uint64_t cap = std::max<uint64_t>(ix->capacity, 1);
  ix->raw.reserve(cap * ix->raw_stride);
  if (ix->params.quantization_type != AB_QUANT_NONE) {
    ix->codes.reserve(cap * ix->code_stride);
    AB_CUDA(cudaMemsetAsync(ix->codes.p, 0, cap * ix->code_stride, ix->stream));
    ix->qmin.reserve(ix->dim);
    ix->qmax.reserve(ix->dim);
    ix->scale.reserve(ix->table_len);
    ix->vmin.reserve(ix->table_len);
  }
  AB_CUDA(cudaMemsetAsync(ix->raw.p, 0, cap * ix->raw_stride * sizeof(float), ix->stream));
  // the reference's Graph storage is filled with -1 (graph.hpp:65-68)
  ix->adj.reserve(cap * 32);
  AB_CUDA(cudaMemsetAsync(ix->adj.p, 0xFF, cap * 32 * sizeof(uint32_t), ix->stream));
  ix->ov_off.reserve(cap);
  AB_CUDA(cudaMemsetAsync(ix->ov_off.p, 0xFF, cap * sizeof(uint32_t), ix->stream));
  ix->h_levels.assign(cap, 0);
  ix->h_valid_raw.assign(cap, 0);
  ix->h_valid_space.assign(cap, 0);
  ix->h_valid_graph.assign(cap, 0);
  ix->has_removed = false;
  ix->inserted_edges.clear();
  ix->removed_vertices.clear();
  ix->removed_node_nbrs.clear();
}

// gmin / gmax: quantiser range to use instead of this shard's own (one logical index split over several GPUs must
// quantise every shard with the global per-dimension range); null = fit on these rows
void fit_spaces(ab_index *ix, const float *data, uint64_t n, uint32_t dim, const float *gmin, const float *gmax) {
  AB_CHECK(data != nullptr && n > 0 && dim > 0, AB_ERR_INVALID, "fit: empty input");
  AB_CHECK(n < (1ull << 31), AB_ERR_INVALID, "fit: ids are 31-bit (query_utils.hpp:302-306)");
  // RawSpace::fit silently drops rows past capacity and leaves a broken graph (sequential_storage.hpp:77-80);
  // SQ spaces throw (sq8_space.hpp:117-119). Both are reported here.
  AB_CHECK(n <= ix->capacity, AB_ERR_CAPACITY, "The number of data points exceeds the capacity of the space");
  use_device(ix);
  index_order_after_searches(ix, ix->stream);
  AB_CUDA(cudaStreamSynchronize(ix->stream));  // buffers may be reallocated below
  index_configure_layout(ix, dim);
  index_alloc_spaces(ix);
  ix->n = ix->n_raw = n;
  // rows may already live on this device (cudaMalloc'd by the caller, e.g. generated there): staged in place
  cudaPointerAttributes at{};
  const bool on_device = cudaPointerGetAttributes(&at, data) == cudaSuccess && at.type == cudaMemoryTypeDevice;
  cudaGetLastError();
  const bool cosine = ix->params.metric == AB_METRIC_COS;
  if (on_device) {
    AB_CHECK(at.device == ix->device, AB_ERR_INVALID, "fit: device rows must live on the index's device");
    const uint64_t chunk_rows = std::max<uint64_t>(1, (1ull << 30) / (size_t(dim) * 4));  // bounded grids
    for (uint64_t r = 0; r < n; r += chunk_rows) {
      const uint64_t m = std::min(chunk_rows, n - r);
      stage_rows(data + r * dim, ix->raw.p + r * ix->raw_stride, m, dim, ix->raw_stride, cosine, ix->stream);
    }
  } else {
    DevBuf<float> staging;
    const uint64_t chunk_rows = std::max<uint64_t>(1, std::min<uint64_t>(n, (256ull << 20) / (size_t(dim) * 4)));
    staging.reserve(chunk_rows * dim);
    for (uint64_t r = 0; r < n; r += chunk_rows) {
      const uint64_t m = std::min(chunk_rows, n - r);
      AB_CUDA(cudaMemcpyAsync(staging.p, data + r * dim, m * dim * sizeof(float), cudaMemcpyHostToDevice, ix->stream));
      stage_rows(staging.p, ix->raw.p + r * ix->raw_stride, m, dim, ix->raw_stride, cosine, ix->stream);
      AB_CUDA(cudaStreamSynchronize(ix->stream));
    }
  }
  const int q = ix->params.quantization_type;
  if (q != AB_QUANT_NONE) {
    // the SQ space is fitted on the rows as stored by the raw space (normalised for COS, because the
    // reference normalised the caller's buffer in place first: index.hpp:197-206)
    if (gmin != nullptr && gmax != nullptr) {
      AB_CUDA(cudaMemcpyAsync(ix->qmin.p, gmin, dim * sizeof(float), cudaMemcpyDefault, ix->stream));
      AB_CUDA(cudaMemcpyAsync(ix->qmax.p, gmax, dim * sizeof(float), cudaMemcpyDefault, ix->stream));
    } else {
      quant_fit(ix->raw.p, n, dim, ix->raw_stride, ix->qmin.p, ix->qmax.p, ix->stream);
    }
    quant_encode(q, ix->raw.p, n, dim, ix->raw_stride, ix->qmin.p, ix->qmax.p, ix->codes.p, ix->code_stride, ix->stream);
    quant_tables(q, ix->qmin.p, ix->qmax.p, dim, ix->scale.p, ix->vmin.p, ix->table_len, ix->stream);
  }
  std::fill(ix->h_valid_raw.begin(), ix->h_valid_raw.begin() + n, 1);
  std::fill(ix->h_valid_space.begin(), ix->h_valid_space.begin() + n, 1);
  ix->has_graph = false;
  AB_CUDA(cudaStreamSynchronize(ix->stream));
}

void index_upload_graph(ab_index *ix, const uint32_t *adj, const uint32_t *levels, const uint32_t *lists_flat,
                        uint64_t lists_len, uint32_t ep) {
  use_device(ix);
  index_order_after_searches(ix, ix->stream);
  const uint64_t n = ix->n;
  AB_CHECK(ep < n, AB_ERR_INVALID, "set_graph: entry point out of range");
  AB_CUDA(cudaMemcpyAsync(ix->adj.p, adj, n * 32 * sizeof(uint32_t), cudaMemcpyHostToDevice, ix->stream));
  std::vector<uint32_t> off(ix->capacity, kEmptyId);
  uint64_t rows = 0;
  ix->has_overlay = levels != nullptr;
  if (levels != nullptr) {
    for (uint64_t i = 0; i < n; i++) {
      ix->h_levels[i] = levels[i];
      if (levels[i] > 0) {
        off[i] = static_cast<uint32_t>(rows);
        rows += levels[i];
      }
    }
    AB_CHECK(rows * 32 == lists_len, AB_ERR_INVALID, "set_graph: overlay list length does not match the levels");
  }
  ix->ov_rows = rows;
  ix->ov_lists.reserve(std::max<uint64_t>(rows, 1) * 32);
  if (rows > 0)
    AB_CUDA(cudaMemcpyAsync(ix->ov_lists.p, lists_flat, rows * 32 * sizeof(uint32_t), cudaMemcpyHostToDevice, ix->stream));
  AB_CUDA(cudaMemcpyAsync(ix->ov_off.p, off.data(), ix->capacity * sizeof(uint32_t), cudaMemcpyHostToDevice, ix->stream));
  ix->ep = ep;
  ix->ep_level = levels != nullptr ? levels[ep] : 0;
  std::fill(ix->h_valid_graph.begin(), ix->h_valid_graph.begin() + n, 1);
  ix->has_graph = true;
  AB_CUDA(cudaStreamSynchronize(ix->stream));
}

uint64_t index_lists_len(const ab_index *ix) { return ix->ov_rows * 32; }

void index_download_graph(ab_index *ix, uint32_t *adj, uint32_t *levels, uint32_t *lists_flat) {
  use_device(ix);
  AB_CHECK(ix->has_graph, AB_ERR_STATE, "index has no graph");
  if (adj) AB_CUDA(cudaMemcpyAsync(adj, ix->adj.p, ix->n * 32 * sizeof(uint32_t), cudaMemcpyDeviceToHost, ix->stream));
  if (levels) std::copy(ix->h_levels.begin(), ix->h_levels.begin() + ix->n, levels);
  if (lists_flat && ix->ov_rows > 0)
    AB_CUDA(cudaMemcpyAsync(lists_flat, ix->ov_lists.p, ix->ov_rows * 32 * sizeof(uint32_t), cudaMemcpyDeviceToHost, ix->stream));
  AB_CUDA(cudaStreamSynchronize(ix->stream));
}

void index_order_after_searches(ab_index *ix, cudaStream_t stream) {
  if (ix->ws_event_recorded) AB_CUDA(cudaStreamWaitEvent(stream, ix->ws_event, 0));
}

void index_mark_workspace_use(ab_index *ix, cudaStream_t stream) {
  if (ix->ws_event == nullptr) AB_CUDA(cudaEventCreateWithFlags(&ix->ws_event, cudaEventDisableTiming));
  AB_CUDA(cudaEventRecord(ix->ws_event, stream));
  ix->ws_event_recorded = true;
}

void index_search_device(ab_index *ix, const float *d_queries, uint64_t nq, uint32_t topk, uint32_t ef, bool rerank,
                         uint32_t *d_out_ids, float *d_out_dist, uint32_t *d_pool_ids, float *d_pool_dist,
                         uint32_t *d_pool_sizes, cudaStream_t stream, const Exchange *exchange, bool pad_empty,
                         const uint32_t *q_ready, uint32_t q_chunk) {
  AB_CHECK(ix->has_graph, AB_ERR_STATE, "index is not fitted or loaded");
  AB_CHECK(topk >= 1 && ef >= 1, AB_ERR_INVALID, "topk and ef must be positive");
  // the reference reads uninitialised pool slots when topk > ef (SURVEY.md 8b "Errors"); rejected here
  AB_CHECK(topk <= ef, AB_ERR_INVALID, "topk must not exceed ef");
  // the pool of one query lives in shared memory (8 bytes per entry): that, not the visited set, bounds ef
  AB_CHECK(ef <= 24576, AB_ERR_INVALID, "ef larger than 24576 is not supported (the candidate pool must fit one SM's shared memory)");
  if (nq == 0) return;
  // the workspace below is shared by every launch on this handle: order this one behind the previous user
  index_order_after_searches(ix, stream);
  const int quant = ix->params.quantization_type;
  const int nchunks = quant == AB_QUANT_NONE ? ix->raw_chunks : ix->code_chunks;
  const bool need_rerank = rerank && quant != AB_QUANT_NONE;
  const bool exch_in_search = exchange != nullptr && !need_rerank;
  AB_CHECK(exchange == nullptr || d_out_dist != nullptr, AB_ERR_INVALID, "the shard exchange needs the distance output");
  SearchPlan plan = plan_search(quant, nchunks, ix->table_len, ix->visit_ratio, ef, static_cast<uint32_t>(nq),
                                static_cast<uint32_t>(ix->n), ix->sm_count, ix->smem_limit,
                                exch_in_search ? size_t(exchange->n_shards) * topk * 8 : 0);
  // global-memory part of the visited set (overflow tables per resident warp + pooled bitmaps)
  ix->w_visited.prepare(plan, std::max<uint64_t>(ix->n, 1), stream);
  ix->w_counter.reserve(1);
  ix->w_stats.reserve(8);
  // feedback from earlier launches (visited ids per query size the visited table), without blocking the stream
  if (ix->stats_pending && cudaEventQuery(ix->stats_event) == cudaSuccess) {
    ix->stats_pending = false;
    if (ix->stats_nq >= 32 && ix->stats_ef > 0)
      ix->visit_ratio = static_cast<float>(double(ix->h_stats[0]) / double(ix->stats_nq) / double(ix->stats_ef)) /
                        visited_model(ix->stats_ef);
  }
  cudaGetLastError();  // cudaErrorNotReady from the query is not an error
  // one measurement per ef value is enough: a steady loop at one ef pays for the copy once
  if (!ix->stats_pending && ix->last_ef > 0 && ix->last_nq >= 32 && ix->last_ef != ix->stats_ef) {
    if (ix->h_stats == nullptr) {
      AB_CUDA(cudaMallocHost(reinterpret_cast<void **>(&ix->h_stats), 4 * sizeof(unsigned long long)));
      AB_CUDA(cudaEventCreateWithFlags(&ix->stats_event, cudaEventDisableTiming));
    }
    AB_CUDA(cudaMemcpyAsync(ix->h_stats, ix->w_stats.p, 4 * sizeof(unsigned long long), cudaMemcpyDeviceToHost, stream));
    AB_CUDA(cudaEventRecord(ix->stats_event, stream));
    ix->stats_pending = true;
    ix->stats_ef = ix->last_ef;
    ix->stats_nq = ix->last_nq;
  }
  AB_CUDA(cudaMemsetAsync(ix->w_counter.p, 0, sizeof(uint32_t), stream));
  AB_CUDA(cudaMemsetAsync(ix->w_stats.p, 0, 8 * sizeof(unsigned long long), stream));

  if (need_rerank) ix->w_cand.reserve(nq * topk);

  SearchArgs a{};
  a.rows = quant == AB_QUANT_NONE ? reinterpret_cast<const uint8_t *>(ix->raw.p) : ix->codes.p;
  a.row_stride = static_cast<uint32_t>(quant == AB_QUANT_NONE ? size_t(ix->raw_stride) * 4 : ix->code_stride);
  a.nchunks = nchunks;
  a.table_len = ix->table_len;
  a.qmin = ix->qmin.p;
  a.qmax = ix->qmax.p;
  a.scale = ix->scale.p;
  a.vmin = ix->vmin.p;
  a.adj = ix->adj.p;
  a.ov_off = ix->has_overlay ? ix->ov_off.p : nullptr;
  a.ov_lists = ix->ov_lists.p;
  a.valid = (quant == AB_QUANT_NONE && ix->has_removed) ? ix->valid.p : nullptr;
  a.ep = ix->ep;
  a.ep_level = ix->ep_level;
  a.n = static_cast<uint32_t>(ix->n);
  a.dim = ix->dim;
  a.metric = ix->params.metric;
  a.queries = d_queries;
  a.q_ids = nullptr;
  a.q_stride = ix->dim;
  a.q_prenormalized = 0;
  a.q_ready = q_ready;
  a.q_chunk = q_chunk;
  a.nq = static_cast<uint32_t>(nq);
  a.ef = ef;
  a.topk = topk;
  a.out_ids = need_rerank ? ix->w_cand.p : d_out_ids;
  a.out_dist = need_rerank ? nullptr : d_out_dist;
  a.pool_ids = d_pool_ids;
  a.pool_dist = d_pool_dist;
  a.pool_sizes = d_pool_sizes;
  a.work_counter = ix->w_counter.p;
  a.hash_size = plan.hash_size;
  ix->w_visited.fill(a);
  a.stats = ix->w_stats.p;
  if (exch_in_search) a.exchange = *exchange;
  a.pad_empty = ((exchange != nullptr && need_rerank) || pad_empty) ? 1 : 0;
  launch_search(quant, plan, a, stream);
  ix->last_ef = ef;
  ix->last_nq = nq;

  if (need_rerank) {
    RerankArgs r{};
    r.raw = ix->raw.p;
    r.raw_stride = size_t(ix->raw_stride) * 4;
    r.nchunks = ix->raw_chunks;
    r.valid = ix->has_removed ? ix->valid.p : nullptr;
    r.n_rows = static_cast<uint32_t>(ix->n_raw);
    r.queries = d_queries;
    r.nq = static_cast<uint32_t>(nq);
    r.dim = ix->dim;
    r.ef = ef;
    r.topk = topk;
    r.metric = ix->params.metric;
    r.cand_ids = ix->w_cand.p;
    r.out_ids = d_out_ids;
    r.out_dist = d_out_dist;
    if (exchange != nullptr) r.exchange = *exchange;
    r.no_zero_fill = (exchange != nullptr || pad_empty) ? 1 : 0;
    launch_rerank(r, ix->sm_count, stream);
  }
  index_mark_workspace_use(ix, stream);
}

// host queries in, host ids out (PyIndex::batch_search shape)
static void search_host(ab_index *ix, const float *queries, uint64_t nq, uint32_t topk, uint32_t ef, bool rerank,
                        uint32_t *out_ids, float *out_dist, uint32_t *pool_ids, float *pool_dist, uint32_t *pool_sizes) {
  std::lock_guard<std::mutex> lock(ix->mu);
  use_device(ix);
  AB_CHECK(ix->has_graph, AB_ERR_STATE, "index is not fitted or loaded");
  if (nq == 0) return;
  AB_CHECK(queries != nullptr && out_ids != nullptr, AB_ERR_INVALID, "null buffer");
  cudaStream_t s = ix->stream;
  // Pinned (page-locked, mapped) caller buffers are used in place: the kernel reads each query once over the
  // host link while other warps compute, and writes the few result bytes straight back, instead of a
  // serialised copy-in / copy-out around the launch. Pageable buffers go through staging copies.
  auto mapped = [](const void *p) -> void * {
    cudaPointerAttributes at{};
    if (p == nullptr || cudaPointerGetAttributes(&at, p) != cudaSuccess) {
      cudaGetLastError();
      return nullptr;
    }
    return at.type == cudaMemoryTypeHost ? at.devicePointer : nullptr;
  };
  const float *d_q = static_cast<const float *>(mapped(queries));
  uint32_t *d_ids = static_cast<uint32_t *>(mapped(out_ids));
  float *d_dist = out_dist ? static_cast<float *>(mapped(out_dist)) : nullptr;
  const bool stage_ids = d_ids == nullptr, stage_dist = out_dist != nullptr && d_dist == nullptr;
  // Pageable queries: a plain cudaMemcpyAsync would serialise "stage + copy" in front of the kernel. For batches
  // worth it the copy runs in chunks on a second stream WHILE the search kernel runs: the kernel is launched first,
  // its warps pull queries in order and wait for the arrival flag of the chunk their query is in, which the copy
  // engine writes right behind the chunk's data.
  constexpr uint32_t kQueryChunks = 8;
  const bool pipelined = d_q == nullptr && nq >= 2048 && pool_ids == nullptr;
  uint32_t q_chunk = 0;
  if (d_q == nullptr) {
    ix->w_queries.reserve(nq * ix->dim);
    if (!pipelined)
      AB_CUDA(cudaMemcpyAsync(ix->w_queries.p, queries, nq * ix->dim * sizeof(float), cudaMemcpyHostToDevice, s));
    d_q = ix->w_queries.p;
  }
  if (pipelined) {
    if (ix->copy_stream == nullptr) {
      AB_CUDA(cudaStreamCreateWithFlags(&ix->copy_stream, cudaStreamNonBlocking));
      AB_CUDA(cudaEventCreateWithFlags(&ix->copy_event, cudaEventDisableTiming));
      AB_CUDA(cudaMallocHost(reinterpret_cast<void **>(&ix->h_one), sizeof(uint32_t)));
      *ix->h_one = 1u;
    }
    ix->w_qready.reserve(kQueryChunks);
    q_chunk = static_cast<uint32_t>((nq + kQueryChunks - 1) / kQueryChunks);
    index_order_after_searches(ix, s);  // the previous kernel may still be polling the old flags
    AB_CUDA(cudaMemsetAsync(ix->w_qready.p, 0, kQueryChunks * sizeof(uint32_t), s));
    AB_CUDA(cudaEventRecord(ix->copy_event, s));
    AB_CUDA(cudaStreamWaitEvent(ix->copy_stream, ix->copy_event, 0));
  }
  if (stage_ids) {
    ix->w_ids.reserve(nq * topk);
    d_ids = ix->w_ids.p;
  }
  if (stage_dist) {
    ix->w_dist.reserve(nq * topk);
    d_dist = ix->w_dist.p;
  }
  if (pool_ids) {
    ix->w_pool_ids.reserve(nq * ef);
    ix->w_pool_dist.reserve(nq * ef);
    ix->w_pool_sizes.reserve(nq);
  }
  index_search_device(ix, d_q, nq, topk, ef, rerank, d_ids, out_dist ? d_dist : nullptr,
                      pool_ids ? ix->w_pool_ids.p : nullptr, pool_ids ? ix->w_pool_dist.p : nullptr,
                      pool_ids ? ix->w_pool_sizes.p : nullptr, s, nullptr, false, pipelined ? ix->w_qready.p : nullptr, q_chunk);
  if (pipelined) {
    // the kernel is in flight; now feed it (each pageable copy returns once its chunk is staged)
    for (uint32_t c = 0; c < kQueryChunks; c++) {
      const uint64_t r0 = uint64_t(c) * q_chunk, r1 = std::min<uint64_t>(nq, r0 + q_chunk);
      if (r0 < r1)
        AB_CUDA(cudaMemcpyAsync(ix->w_queries.p + r0 * ix->dim, queries + r0 * ix->dim, (r1 - r0) * ix->dim * sizeof(float),
                                cudaMemcpyHostToDevice, ix->copy_stream));
      AB_CUDA(cudaMemcpyAsync(ix->w_qready.p + c, ix->h_one, sizeof(uint32_t), cudaMemcpyHostToDevice, ix->copy_stream));
    }
  }
  if (stage_ids) AB_CUDA(cudaMemcpyAsync(out_ids, ix->w_ids.p, nq * topk * sizeof(uint32_t), cudaMemcpyDeviceToHost, s));
  if (stage_dist) AB_CUDA(cudaMemcpyAsync(out_dist, ix->w_dist.p, nq * topk * sizeof(float), cudaMemcpyDeviceToHost, s));
  if (pool_ids) {
    AB_CUDA(cudaMemcpyAsync(pool_ids, ix->w_pool_ids.p, nq * ef * sizeof(uint32_t), cudaMemcpyDeviceToHost, s));
    AB_CUDA(cudaMemcpyAsync(pool_dist, ix->w_pool_dist.p, nq * ef * sizeof(float), cudaMemcpyDeviceToHost, s));
    AB_CUDA(cudaMemcpyAsync(pool_sizes, ix->w_pool_sizes.p, nq * sizeof(uint32_t), cudaMemcpyDeviceToHost, s));
  }
  unsigned long long st[6];
  AB_CUDA(cudaMemcpyAsync(st, ix->w_stats.p, sizeof(st), cudaMemcpyDeviceToHost, s));
  AB_CUDA(cudaStreamSynchronize(s));
  for (int i = 0; i < 4; i++) ix->last_stats[i] = st[i];
  ix->last_stats[4] = nq;
  ix->last_stats[5] = st[4];
  ix->last_stats[6] = st[5];
  if (nq >= 32 && ix->last_ef > 0)
    ix->visit_ratio = static_cast<float>(double(st[0]) / double(nq) / double(ix->last_ef)) / visited_model(ix->last_ef);
}

}  // namespace ab

using namespace ab;  // NOLINT

extern "C" {

const char *ab_last_error(void) { return g_last_error.c_str(); }
const char *ab_version(void) { return "alayalite_b200 0.1 (sm_100a)"; }
uint64_t ab_kernel_launch_count(void) { return g_kernel_launches.load(); }

int ab_device_count(void) {
  int n = 0;
  if (cudaGetDeviceCount(&n) != cudaSuccess) {
    cudaGetLastError();
    return 0;
  }
  return n;
}
int ab_set_device(int32_t device) {
  return guarded([&] { AB_CUDA(cudaSetDevice(device)); });
}

int ab_index_create(const ab_params *params, ab_index **out) {
  return guarded([&] {
    AB_CHECK(params != nullptr && out != nullptr, AB_ERR_INVALID, "null argument");
    AB_CHECK(params->index_type == AB_INDEX_HNSW || params->index_type == AB_INDEX_NSG ||
                 params->index_type == AB_INDEX_FUSION,
             AB_ERR_UNSUPPORTED, "Unsupported index type");  // dispatch.hpp:119-121
    AB_CHECK(params->quantization_type >= AB_QUANT_NONE && params->quantization_type <= AB_QUANT_SQ4,
             AB_ERR_UNSUPPORTED, "Unsupported quantization type");
    AB_CHECK(params->metric >= AB_METRIC_L2 && params->metric <= AB_METRIC_COS, AB_ERR_UNSUPPORTED, "Unsupported metric");
    AB_CHECK(params->capacity > 0, AB_ERR_INVALID, "Capacity must be greater than 0");
    AB_CHECK(params->max_nbrs == 0 || params->max_nbrs == 32, AB_ERR_UNSUPPORTED,
             "max_nbrs is fixed at 32 (the reference never forwards another value, schema.py:94-101)");
    AB_CHECK(ab_device_count() > 0, AB_ERR_CUDA, "no CUDA device: libalaya_b200 has no CPU fallback");
    auto ix = std::make_unique<ab_index>();
    ix->params = *params;
    ix->params.max_nbrs = 32;
    ix->capacity = params->capacity;
    int dev = params->device;
    if (dev < 0) AB_CUDA(cudaGetDevice(&dev));
    ix->device = dev;
    AB_CUDA(cudaSetDevice(dev));
    cudaDeviceProp prop{};
    AB_CUDA(cudaGetDeviceProperties(&prop, dev));
    AB_CHECK(prop.major == 10, AB_ERR_CUDA, "libalaya_b200 is built for sm_100a (B200) only");
    ix->sm_count = prop.multiProcessorCount;
    ix->smem_limit = prop.sharedMemPerBlockOptin;
    AB_CUDA(cudaStreamCreateWithFlags(&ix->stream, cudaStreamNonBlocking));
    *out = ix.release();
  });
}

void ab_index_destroy(ab_index *index) {
  if (index == nullptr) return;
  cudaSetDevice(index->device);
  if (index->stream) {
    cudaStreamSynchronize(index->stream);
    cudaStreamDestroy(index->stream);
  }
  if (index->stats_event) {
    cudaEventSynchronize(index->stats_event);
    cudaEventDestroy(index->stats_event);
  }
  if (index->ws_event) {
    cudaEventSynchronize(index->ws_event);
    cudaEventDestroy(index->ws_event);
  }
  if (index->copy_stream) {
    cudaStreamSynchronize(index->copy_stream);
    cudaStreamDestroy(index->copy_stream);
    cudaEventDestroy(index->copy_event);
    cudaFreeHost(index->h_one);
  }
  if (index->h_stats) cudaFreeHost(index->h_stats);
  delete index;
}

int ab_index_fit_spaces(ab_index *index, const float *data, uint64_t n, uint32_t dim) {
  return guarded([&] {
    AB_CHECK(index != nullptr, AB_ERR_INVALID, "null index");
    std::lock_guard<std::mutex> lock(index->mu);
    fit_spaces(index, data, n, dim, nullptr, nullptr);
  });
}

int ab_index_build_graph(ab_index *index, uint32_t ef_construction) {
  return guarded([&] {
    AB_CHECK(index != nullptr && index->n > 0, AB_ERR_STATE, "fit the spaces first");
    std::lock_guard<std::mutex> lock(index->mu);
    use_device(index);
    index_order_after_searches(index, index->stream);
    index_build_graph(index, ef_construction);
  });
}

int ab_index_build_graph_nsg(ab_index *index, uint32_t ef_construction, uint32_t nndescent_iters) {
  return guarded([&] {
    AB_CHECK(index != nullptr && index->n > 0, AB_ERR_STATE, "fit the spaces first");
    std::lock_guard<std::mutex> lock(index->mu);
    use_device(index);
    index_order_after_searches(index, index->stream);
    index_build_graph_nsg(index, ef_construction, nndescent_iters == 0 ? 10 : nndescent_iters);
  });
}

int ab_index_knng_nndescent(ab_index *index, uint32_t k, uint32_t iters, uint32_t *out_ids) {
  return guarded([&] {
    AB_CHECK(index != nullptr && index->n > 0 && out_ids != nullptr, AB_ERR_STATE, "fit the spaces first");
    std::lock_guard<std::mutex> lock(index->mu);
    use_device(index);
    index_order_after_searches(index, index->stream);
    index_knng_nndescent(index, k, iters == 0 ? 10 : iters, out_ids);
  });
}

int ab_index_fit(ab_index *index, const float *data, uint64_t n, uint32_t dim, uint32_t ef_construction,
                 uint32_t num_threads) {
  (void)num_threads;
  return guarded([&] {
    AB_CHECK(index != nullptr, AB_ERR_INVALID, "null index");
    std::lock_guard<std::mutex> lock(index->mu);
    fit_spaces(index, data, n, dim, nullptr, nullptr);
    index_build_graph(index, ef_construction);
  });
}

int ab_index_fit_ranged(ab_index *index, const float *data, uint64_t n, uint32_t dim, uint32_t ef_construction, const float *mn,
                        const float *mx) {
  return guarded([&] {
    AB_CHECK(index != nullptr, AB_ERR_INVALID, "null index");
    AB_CHECK((mn == nullptr) == (mx == nullptr), AB_ERR_INVALID, "min and max go together");
    std::lock_guard<std::mutex> lock(index->mu);
    fit_spaces(index, data, n, dim, mn, mx);
    index_build_graph(index, ef_construction);
  });
}

int ab_index_set_graph(ab_index *index, const uint32_t *adj, const uint32_t *levels, const uint32_t *lists_flat,
                       uint64_t lists_len, uint32_t ep) {
  return guarded([&] {
    AB_CHECK(index != nullptr && index->n > 0, AB_ERR_STATE, "fit the spaces first");
    std::lock_guard<std::mutex> lock(index->mu);
    index_upload_graph(index, adj, levels, lists_flat, lists_len, ep);
  });
}

int ab_index_graph_info(ab_index *index, uint64_t info[5]) {
  return guarded([&] {
    AB_CHECK(index != nullptr, AB_ERR_INVALID, "null index");
    info[0] = index->n;
    info[1] = 32;
    info[2] = index->ep;
    info[3] = index_lists_len(index);
    info[4] = index->capacity;
  });
}

int ab_index_get_graph(ab_index *index, uint32_t *adj, uint32_t *levels, uint32_t *lists_flat) {
  return guarded([&] {
    AB_CHECK(index != nullptr, AB_ERR_INVALID, "null index");
    std::lock_guard<std::mutex> lock(index->mu);
    index_download_graph(index, adj, levels, lists_flat);
  });
}

int ab_index_get_codes(ab_index *index, uint8_t *codes, float *mn, float *mx) {
  return guarded([&] {
    AB_CHECK(index != nullptr && index->params.quantization_type != AB_QUANT_NONE, AB_ERR_STATE, "index has no SQ space");
    std::lock_guard<std::mutex> lock(index->mu);
    use_device(index);
    if (codes)
      AB_CUDA(cudaMemcpy2DAsync(codes, index->row_bytes, index->codes.p, index->code_stride, index->row_bytes, index->n,
                                cudaMemcpyDeviceToHost, index->stream));
    if (mn) AB_CUDA(cudaMemcpyAsync(mn, index->qmin.p, index->dim * 4, cudaMemcpyDeviceToHost, index->stream));
    if (mx) AB_CUDA(cudaMemcpyAsync(mx, index->qmax.p, index->dim * 4, cudaMemcpyDeviceToHost, index->stream));
    AB_CUDA(cudaStreamSynchronize(index->stream));
  });
}

int ab_index_get_data_by_id(ab_index *index, uint32_t id, float *out) {
  return guarded([&] {
    AB_CHECK(index != nullptr && index->n_raw > 0, AB_ERR_STATE, "space is nullptr");  // index.hpp:99-101
    AB_CHECK(id < index->n_raw, AB_ERR_INVALID, "id out of range");                     // index.hpp:103-105
    std::lock_guard<std::mutex> lock(index->mu);
    use_device(index);
    AB_CUDA(cudaMemcpyAsync(out, index->raw.p + size_t(id) * index->raw_stride, index->dim * sizeof(float),
                            cudaMemcpyDeviceToHost, index->stream));
    AB_CUDA(cudaStreamSynchronize(index->stream));
  });
}

uint32_t ab_index_get_dim(const ab_index *index) { return index ? index->dim : 0; }
uint64_t ab_index_get_size(const ab_index *index) { return index ? index->n : 0; }

int ab_index_search(ab_index *index, const float *query, uint32_t topk, uint32_t ef, uint32_t *out_ids) {
  return guarded([&] {
    AB_CHECK(index != nullptr, AB_ERR_INVALID, "null index");
    search_host(index, query, 1, topk, ef, true, out_ids, nullptr, nullptr, nullptr, nullptr);
  });
}

int ab_index_batch_search(ab_index *index, const float *queries, uint64_t nq, uint32_t topk, uint32_t ef,
                          uint32_t num_threads, uint32_t *out_ids) {
  (void)num_threads;
  return guarded([&] {
    AB_CHECK(index != nullptr, AB_ERR_INVALID, "null index");
    search_host(index, queries, nq, topk, ef, true, out_ids, nullptr, nullptr, nullptr, nullptr);
  });
}

int ab_index_batch_search_with_distance(ab_index *index, const float *queries, uint64_t nq, uint32_t topk, uint32_t ef,
                                        uint32_t num_threads, uint32_t *out_ids, float *out_dist, int32_t *has_dist) {
  (void)num_threads;
  return guarded([&] {
    AB_CHECK(index != nullptr, AB_ERR_INVALID, "null index");
    const bool raw = index->params.quantization_type == AB_QUANT_NONE;
    if (has_dist) *has_dist = raw ? 1 : 0;
    // index.hpp:395-418: SQ spaces return the un-reranked ids and no distances
    search_host(index, queries, nq, topk, ef, false, out_ids, raw ? out_dist : nullptr, nullptr, nullptr, nullptr);
  });
}

int ab_index_batch_traverse(ab_index *index, const float *queries, uint64_t nq, uint32_t ef, uint32_t *pool_ids,
                            float *pool_dist, uint32_t *pool_sizes) {
  return guarded([&] {
    AB_CHECK(index != nullptr && pool_ids && pool_dist && pool_sizes, AB_ERR_INVALID, "null argument");
    std::vector<uint32_t> first(nq);
    search_host(index, queries, nq, 1, ef, false, first.data(), nullptr, pool_ids, pool_dist, pool_sizes);
  });
}

int ab_index_batch_search_device(ab_index *index, const float *d_queries, uint64_t nq, uint32_t topk, uint32_t ef,
                                 int32_t rerank, uint32_t *d_out_ids, float *d_out_dist, void *stream) {
  return guarded([&] {
    AB_CHECK(index != nullptr && d_queries && d_out_ids, AB_ERR_INVALID, "null argument");
    std::lock_guard<std::mutex> lock(index->mu);
    use_device(index);
    cudaStream_t s = static_cast<cudaStream_t>(stream);  // NULL = the legacy default stream, ordered with the caller's work
    index_search_device(index, d_queries, nq, topk, ef, (rerank & 1) != 0, d_out_ids, d_out_dist, nullptr, nullptr, nullptr, s, nullptr,
                        (rerank & 2) != 0);
  });
}

int ab_debug_search_plan(int32_t quant, uint32_t row_chunks, uint32_t ef, uint64_t nq, uint64_t n_rows, int32_t sm_count,
                         uint64_t smem_limit, uint64_t out[6]) {
  return guarded([&] {
    AB_CHECK(out != nullptr && row_chunks >= 1 && ef >= 1 && sm_count >= 1 && smem_limit >= 48 * 1024, AB_ERR_INVALID,
             "bad plan query");
    const int table_len = quant == AB_QUANT_SQ8 ? row_chunks * 16 : (quant == AB_QUANT_SQ4 ? row_chunks * 32 : 0);
    const float vr = getenv("AB_DEBUG_VISIT_RATIO") ? static_cast<float>(atof(getenv("AB_DEBUG_VISIT_RATIO"))) : 0.f;  // developer knob
    const SearchPlan p = plan_search(quant, static_cast<int>(row_chunks), table_len, vr, ef, static_cast<uint32_t>(nq),
                                     static_cast<uint32_t>(n_rows), sm_count, smem_limit);
    out[0] = p.warps_per_cta;
    out[1] = p.ctas_per_sm;
    out[2] = p.ctas;
    out[3] = p.hash_size;
    out[4] = p.hash_rbits != 0 ? 2 : 4;
    out[5] = p.smem_bytes;
  });
}

static void read_stats(ab_index *index) {
  std::lock_guard<std::mutex> lock(index->mu);
  use_device(index);
  if (index->w_stats.p) {
    unsigned long long st[6];
    if (index->ws_event_recorded) AB_CUDA(cudaEventSynchronize(index->ws_event));
    AB_CUDA(cudaMemcpy(st, index->w_stats.p, sizeof(st), cudaMemcpyDeviceToHost));
    for (int i = 0; i < 4; i++) index->last_stats[i] = st[i];
    index->last_stats[5] = st[4];
    index->last_stats[6] = st[5];
  }
}

int ab_index_search_stats(ab_index *index, uint64_t stats[5]) {
  return guarded([&] {
    AB_CHECK(index != nullptr, AB_ERR_INVALID, "null index");
    read_stats(index);
    for (int i = 0; i < 5; i++) stats[i] = index->last_stats[i];
  });
}

int ab_index_search_stats_ex(ab_index *index, uint64_t *stats, uint32_t count) {
  return guarded([&] {
    AB_CHECK(index != nullptr && stats != nullptr, AB_ERR_INVALID, "null argument");
    read_stats(index);
    for (uint32_t i = 0; i < count; i++) stats[i] = i < 7 ? index->last_stats[i] : 0;
    if (count > 7) stats[7] = index->w_visited.bytes();
  });
}

int ab_index_insert(ab_index *index, const float *vec, uint32_t ef, uint32_t *out_id) {
  return guarded([&] {
    AB_CHECK(index != nullptr && vec && out_id, AB_ERR_INVALID, "null argument");
    *out_id = index_insert(index, vec, ef);
  });
}
int ab_index_remove(ab_index *index, uint32_t id) {
  return guarded([&] {
    AB_CHECK(index != nullptr, AB_ERR_INVALID, "null index");
    index_remove(index, id);
  });
}

int ab_index_save(ab_index *index, const char *index_path, const char *data_path, const char *quant_path) {
  return guarded([&] {
    AB_CHECK(index != nullptr && index_path, AB_ERR_INVALID, "null argument");
    std::lock_guard<std::mutex> lock(index->mu);
    use_device(index);
    index_order_after_searches(index, index->stream);
    index_save(index, index_path, data_path, quant_path);
  });
}
int ab_index_load(ab_index *index, const char *index_path, const char *data_path, const char *quant_path) {
  return guarded([&] {
    AB_CHECK(index != nullptr && index_path, AB_ERR_INVALID, "null argument");
    std::lock_guard<std::mutex> lock(index->mu);
    use_device(index);
    index_order_after_searches(index, index->stream);
    AB_CUDA(cudaStreamSynchronize(index->stream));
    index_load(index, index_path, data_path, quant_path);
  });
}

// ---- stand-alone kernels -----------------------------------------------------------------------
static void require_device() { AB_CHECK(ab_device_count() > 0, AB_ERR_CUDA, "no CUDA device: libalaya_b200 has no CPU fallback"); }

int ab_distance_pairs(int32_t metric, int32_t quant, const void *x, const void *y, uint64_t pairs, uint32_t dim,
                      const float *mn, const float *mx, float *out) {
  return guarded([&] {
    require_device();
    AB_CHECK(x && y && out && dim > 0, AB_ERR_INVALID, "null argument");
    if (pairs == 0) return;
    ab_index lay;
    lay.params.quantization_type = quant;
    index_configure_layout(&lay, dim);
    const bool raw = quant == AB_QUANT_NONE;
    AB_CHECK(raw || (mn && mx), AB_ERR_INVALID, "SQ distances need min/max");
    const size_t in_row = raw ? size_t(dim) * 4 : lay.row_bytes;
    const size_t stride = raw ? size_t(lay.raw_stride) * 4 : lay.code_stride;
    const int nchunks = raw ? lay.raw_chunks : lay.code_chunks;
    cudaStream_t s = nullptr;
    DevBuf<uint8_t> dx, dy;
    DevBuf<float> dout, dmn, dmx, dscale, dvmin;
    dx.alloc_zero(pairs * stride, s);
    dy.alloc_zero(pairs * stride, s);
    dout.reserve(pairs);
    AB_CUDA(cudaMemcpy2DAsync(dx.p, stride, x, in_row, in_row, pairs, cudaMemcpyHostToDevice, s));
    AB_CUDA(cudaMemcpy2DAsync(dy.p, stride, y, in_row, in_row, pairs, cudaMemcpyHostToDevice, s));
    if (!raw) {
      dmn.reserve(dim);
      dmx.reserve(dim);
      dscale.reserve(lay.table_len);
      dvmin.reserve(lay.table_len);
      AB_CUDA(cudaMemcpyAsync(dmn.p, mn, dim * 4, cudaMemcpyHostToDevice, s));
      AB_CUDA(cudaMemcpyAsync(dmx.p, mx, dim * 4, cudaMemcpyHostToDevice, s));
      quant_tables(quant, dmn.p, dmx.p, dim, dscale.p, dvmin.p, lay.table_len, s);
    }
    distance_pairs(metric, quant, dx.p, dy.p, stride, nchunks, pairs, dscale.p, dvmin.p, lay.table_len, dout.p, s);
    AB_CUDA(cudaMemcpyAsync(out, dout.p, pairs * 4, cudaMemcpyDeviceToHost, s));
    AB_CUDA(cudaStreamSynchronize(s));
  });
}

int ab_quant_fit_encode(int32_t quant, const float *data, uint64_t n, uint32_t dim, float *mn, float *mx, uint8_t *codes) {
  return guarded([&] {
    require_device();
    AB_CHECK(quant == AB_QUANT_SQ8 || quant == AB_QUANT_SQ4, AB_ERR_UNSUPPORTED, "Unsupported quantization type");
    AB_CHECK(data && n > 0 && dim > 0, AB_ERR_INVALID, "empty input");
    ab_index lay;
    lay.params.quantization_type = quant;
    index_configure_layout(&lay, dim);
    cudaStream_t s = nullptr;
    DevBuf<float> d, dmn, dmx;
    DevBuf<uint8_t> dc;
    d.reserve(n * dim);
    dmn.reserve(dim);
    dmx.reserve(dim);
    dc.reserve(n * lay.row_bytes);
    AB_CUDA(cudaMemcpyAsync(d.p, data, n * dim * 4, cudaMemcpyHostToDevice, s));
    quant_fit(d.p, n, dim, dim, dmn.p, dmx.p, s);
    quant_encode(quant, d.p, n, dim, dim, dmn.p, dmx.p, dc.p, lay.row_bytes, s);
    if (mn) AB_CUDA(cudaMemcpyAsync(mn, dmn.p, dim * 4, cudaMemcpyDeviceToHost, s));
    if (mx) AB_CUDA(cudaMemcpyAsync(mx, dmx.p, dim * 4, cudaMemcpyDeviceToHost, s));
    if (codes) AB_CUDA(cudaMemcpyAsync(codes, dc.p, n * lay.row_bytes, cudaMemcpyDeviceToHost, s));
    AB_CUDA(cudaStreamSynchronize(s));
  });
}

int ab_quant_minmax(int32_t metric, const float *data, uint64_t n, uint32_t dim, float *mn, float *mx) {
  return guarded([&] {
    require_device();
    AB_CHECK(data && mn && mx && n > 0 && dim > 0, AB_ERR_INVALID, "empty input");
    cudaStream_t s = nullptr;
    cudaPointerAttributes at{};
    const bool on_device = cudaPointerGetAttributes(&at, data) == cudaSuccess && at.type == cudaMemoryTypeDevice;
    cudaGetLastError();
    const uint64_t chunk = std::max<uint64_t>(1, std::min<uint64_t>(n, (256ull << 20) / (size_t(dim) * 4)));
    DevBuf<float> a, b, dmn, dmx;
    if (!on_device) a.reserve(chunk * dim);
    b.reserve(chunk * dim);
    dmn.reserve(dim);
    dmx.reserve(dim);
    std::vector<float> cmn(dim), cmx(dim);
    std::fill(mn, mn + dim, FLT_MAX);
    std::fill(mx, mx + dim, -FLT_MAX);
    for (uint64_t r = 0; r < n; r += chunk) {
      const uint64_t m = std::min(chunk, n - r);
      const float *src = data + r * dim;
      if (!on_device) {
        AB_CUDA(cudaMemcpyAsync(a.p, src, m * dim * 4, cudaMemcpyHostToDevice, s));
        src = a.p;
      }
      // the SQ space is fitted on the rows as the raw space stores them: normalised for COS (index.hpp:197-206)
      const float *rows = src;
      if (metric == AB_METRIC_COS) {
        stage_rows(src, b.p, m, dim, dim, true, s);
        rows = b.p;
      }
      quant_fit(rows, m, dim, dim, dmn.p, dmx.p, s);
      AB_CUDA(cudaMemcpyAsync(cmn.data(), dmn.p, dim * 4, cudaMemcpyDeviceToHost, s));
      AB_CUDA(cudaMemcpyAsync(cmx.data(), dmx.p, dim * 4, cudaMemcpyDeviceToHost, s));
      AB_CUDA(cudaStreamSynchronize(s));
      for (uint32_t d = 0; d < dim; d++) {
        mn[d] = std::min(mn[d], cmn[d]);
        mx[d] = std::max(mx[d], cmx[d]);
      }
    }
  });
}

int ab_normalize_rows(float *data, uint64_t n, uint32_t dim) {
  return guarded([&] {
    require_device();
    AB_CHECK(data && dim > 0, AB_ERR_INVALID, "null argument");
    if (n == 0) return;
    cudaStream_t s = nullptr;
    DevBuf<float> a, b;
    a.reserve(n * dim);
    b.reserve(n * dim);
    AB_CUDA(cudaMemcpyAsync(a.p, data, n * dim * 4, cudaMemcpyHostToDevice, s));
    stage_rows(a.p, b.p, n, dim, dim, true, s);
    AB_CUDA(cudaMemcpyAsync(data, b.p, n * dim * 4, cudaMemcpyDeviceToHost, s));
    AB_CUDA(cudaStreamSynchronize(s));
  });
}

static void bruteforce_device_impl(int32_t metric, const float *d_data, uint64_t n, uint32_t dim, const float *d_queries,
                                   uint64_t nq, uint32_t k, uint32_t *d_out_ids, float *d_out_dist, cudaStream_t s) {
  int dev = 0, sms = 148;
  AB_CUDA(cudaGetDevice(&dev));
  AB_CUDA(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
  const uint32_t stride = static_cast<uint32_t>(round_up(dim, 16));
  DevBuf<float> pr, pq;
  const float *rows = d_data, *qs = d_queries;
  const bool cosine = metric == AB_METRIC_COS;
  if (stride != dim || cosine) {  // pad (and for COS normalise) into private copies
    pr.reserve(n * stride);
    pq.reserve(nq * stride);
    stage_rows(d_data, pr.p, n, dim, stride, cosine, s);
    stage_rows(d_queries, pq.p, nq, dim, stride, cosine, s);
    rows = pr.p;
    qs = pq.p;
  }
  std::lock_guard<std::mutex> bf_lock(bruteforce_mutex());  // persistent workspaces of both paths
  if (bruteforce_tc_applicable(n, nq, k) && n >= k) {
    bruteforce_topk_tc(metric, rows, n, dim, stride, qs, nq, stride, k, d_out_ids, d_out_dist, sms, s);
  } else {
    bruteforce_topk(metric, rows, n, stride, qs, nq, stride, k, d_out_ids, d_out_dist, sms, s);
  }
  AB_CUDA(cudaStreamSynchronize(s));
}

int ab_bruteforce_topk_device(int32_t metric, const float *d_data, uint64_t n, uint32_t dim, const float *d_queries,
                              uint64_t nq, uint32_t k, uint32_t *d_out_ids, float *d_out_dist, void *stream) {
  return guarded([&] {
    require_device();
    AB_CHECK(d_data && d_queries && d_out_ids && n > 0 && dim > 0, AB_ERR_INVALID, "null argument");
    bruteforce_device_impl(metric, d_data, n, dim, d_queries, nq, k, d_out_ids, d_out_dist, static_cast<cudaStream_t>(stream));
  });
}

int ab_bruteforce_topk(int32_t metric, const float *data, uint64_t n, uint32_t dim, const float *queries, uint64_t nq,
                       uint32_t k, uint32_t *out_ids, float *out_dist) {
  return guarded([&] {
    require_device();
    AB_CHECK(data && queries && out_ids && n > 0 && dim > 0, AB_ERR_INVALID, "null argument");
    if (nq == 0) return;
    cudaStream_t s = nullptr;
    DevBuf<float> dd, dq, dod;
    DevBuf<uint32_t> doi;
    dd.reserve(n * dim);
    dq.reserve(nq * dim);
    doi.reserve(nq * k);
    dod.reserve(nq * k);
    AB_CUDA(cudaMemcpyAsync(dd.p, data, n * dim * 4, cudaMemcpyHostToDevice, s));
    AB_CUDA(cudaMemcpyAsync(dq.p, queries, nq * dim * 4, cudaMemcpyHostToDevice, s));
    bruteforce_device_impl(metric, dd.p, n, dim, dq.p, nq, k, doi.p, dod.p, s);
    AB_CUDA(cudaMemcpyAsync(out_ids, doi.p, nq * k * 4, cudaMemcpyDeviceToHost, s));
    if (out_dist) AB_CUDA(cudaMemcpyAsync(out_dist, dod.p, nq * k * 4, cudaMemcpyDeviceToHost, s));
    AB_CUDA(cudaStreamSynchronize(s));
  });
}

int ab_merge_topk_device(const float *d_dist, const uint32_t *d_ids, uint32_t lists, uint64_t nq, uint32_t k,
                         uint32_t *d_out_ids, float *d_out_dist, void *stream) {
  return guarded([&] {
    require_device();
    AB_CHECK(d_dist && d_ids && d_out_ids && lists > 0 && k > 0, AB_ERR_INVALID, "null argument");
    int dev = 0, sms = 148;
    AB_CUDA(cudaGetDevice(&dev));
    AB_CUDA(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
    merge_topk(d_dist, d_ids, lists, nq, k, d_out_ids, d_out_dist, sms, static_cast<cudaStream_t>(stream));
  });
}

}  // extern "C"
