// Incremental insert / remove.
//
// Replaces (paths relative to /root/reference):
//   GraphUpdateJob::insert_and_update   include/executor/jobs/graph_update_job.hpp:66-89
//   GraphUpdateJob::update              graph_update_job.hpp:105-137
//   GraphUpdateJob::remove              graph_update_job.hpp:91-103
//   JobContext                          include/executor/jobs/job_context.hpp:25-29
//
// The reference does this on one thread with hash maps; it is bookkeeping around two pieces of
// arithmetic, both of which run on the device here: the top-32 search for the new vector (the
// regular traversal kernel) and the distances from each touched node to its candidate neighbours
// (node_distances_kernel). The host only moves ids around, exactly as the reference does, including
// its quirks: the new row goes into the SEARCH space only (:75), unfilled edge slots are written as
// 0 rather than -1 (:132-136), and slots of the search result past the pool size are id 0.
#include <algorithm>
#include <cfloat>

#include "distance.cuh"
#include "index.cuh"
#include "kernels.cuh"

namespace ab {

namespace {

// one warp per touched node: stage its row as the query, stream its candidates
template <int QUANT, bool IP>
__global__ void __launch_bounds__(128) node_distances_kernel(const uint8_t *__restrict__ rows, size_t stride, int nchunks,
                                                             const float *__restrict__ scale, const float *__restrict__ vmin,
                                                             int table_len, const uint32_t *__restrict__ nodes,
                                                             const uint32_t *__restrict__ offsets,
                                                             const uint32_t *__restrict__ cands, uint32_t m,
                                                             float *__restrict__ out) {
  extern __shared__ __align__(16) uint8_t smem[];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, warps = blockDim.x >> 5;
  float *s_scale = reinterpret_cast<float *>(smem);
  float *s_vmin = s_scale + table_len;
  uint8_t *q = smem + size_t(table_len) * 8 + size_t(warp) * nchunks * 16;
  for (int i = threadIdx.x; i < table_len; i += blockDim.x) {
    s_scale[i] = scale[i];
    s_vmin[i] = vmin[i];
  }
  __syncthreads();
  QuerySmem qs{reinterpret_cast<const float *>(q), q, s_scale, s_vmin};
  for (uint32_t t = blockIdx.x * warps + warp; t < m; t += gridDim.x * warps) {
    const uint4 *src = reinterpret_cast<const uint4 *>(rows + size_t(nodes[t]) * stride);
    for (int c = lane; c < nchunks; c += 32) reinterpret_cast<uint4 *>(q)[c] = src[c];
    __syncwarp();
    for (uint32_t e = offsets[t]; e < offsets[t + 1]; e++) {
      const float d = row_distance<QUANT, IP, 32, 4>(rows + size_t(cands[e]) * stride, nchunks, lane, qs, true);
      if (lane == 0) out[e] = d;
    }
    __syncwarp();
  }
}

void node_distances(ab_index *ix, const std::vector<uint32_t> &nodes, const std::vector<uint32_t> &offsets,
                    const std::vector<uint32_t> &cands, std::vector<float> &out) {
  cudaStream_t s = ix->stream;
  const int quant = ix->params.quantization_type;
  const bool raw = quant == AB_QUANT_NONE;
  const bool ip = ix->params.metric != AB_METRIC_L2;
  const uint32_t m = static_cast<uint32_t>(nodes.size());
  out.assign(cands.size(), 0.f);
  if (m == 0 || cands.empty()) return;
  DevBuf<uint32_t> dn, doff, dc;
  DevBuf<float> dout;
  dn.reserve(m);
  doff.reserve(m + 1);
  dc.reserve(cands.size());
  dout.reserve(cands.size());
  AB_CUDA(cudaMemcpyAsync(dn.p, nodes.data(), m * 4, cudaMemcpyHostToDevice, s));
  AB_CUDA(cudaMemcpyAsync(doff.p, offsets.data(), (m + 1) * 4, cudaMemcpyHostToDevice, s));
  AB_CUDA(cudaMemcpyAsync(dc.p, cands.data(), cands.size() * 4, cudaMemcpyHostToDevice, s));
  const uint8_t *rows = raw ? reinterpret_cast<const uint8_t *>(ix->raw.p) : ix->codes.p;
  const size_t stride = raw ? size_t(ix->raw_stride) * 4 : ix->code_stride;
  const int nchunks = raw ? ix->raw_chunks : ix->code_chunks;
  const int warps = 4;
  const size_t smem = size_t(ix->table_len) * 8 + size_t(warps) * nchunks * 16;
  const int blocks = std::max(1u, (m + warps - 1) / warps);
#define AB_ND(Q, IPV)                                                                                      \
  {                                                                                                        \
    auto k = node_distances_kernel<Q, IPV>;                                                                \
    AB_CUDA(cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem))); \
    k<<<blocks, warps * 32, smem, s>>>(rows, stride, nchunks, ix->scale.p, ix->vmin.p, ix->table_len, dn.p, doff.p, dc.p, m, dout.p); \
  }
  if (raw) {
    if (ip) AB_ND(AB_QUANT_NONE, true) else AB_ND(AB_QUANT_NONE, false)
  } else if (quant == AB_QUANT_SQ8) {
    if (ip) AB_ND(AB_QUANT_SQ8, true) else AB_ND(AB_QUANT_SQ8, false)
  } else {
    if (ip) AB_ND(AB_QUANT_SQ4, true) else AB_ND(AB_QUANT_SQ4, false)
  }
#undef AB_ND
  AB_CUDA(cudaGetLastError());
  count_launch();
  AB_CUDA(cudaMemcpyAsync(out.data(), dout.p, cands.size() * 4, cudaMemcpyDeviceToHost, s));
  AB_CUDA(cudaStreamSynchronize(s));
}

}  // namespace

uint32_t index_insert(ab_index *ix, const float *vec, uint32_t ef) {
  std::lock_guard<std::mutex> lock(ix->mu);
  AB_CUDA(cudaSetDevice(ix->device));
  AB_CHECK(ix->has_graph, AB_ERR_STATE, "Index is not init yet");
  // search_solo(query, k = max_nbrs, ...) reads 32 pool slots (graph_update_job.hpp:69); with ef < 32 the
  // reference runs off the end of its pool
  AB_CHECK(ef >= 32, AB_ERR_INVALID, "insert needs ef >= 32 (the reference reads max_nbrs pool slots)");
  cudaStream_t s = ix->stream;
  const int quant = ix->params.quantization_type;
  const bool raw = quant == AB_QUANT_NONE;

  // 1. top-32 of the search space for the new vector (un-reranked pool order)
  std::vector<uint32_t> res(32, kEmptyId);
  ix->w_queries.reserve(ix->dim);
  ix->w_ids.reserve(32);
  AB_CUDA(cudaMemcpyAsync(ix->w_queries.p, vec, ix->dim * sizeof(float), cudaMemcpyHostToDevice, s));
  index_search_device(ix, ix->w_queries.p, 1, 32, ef, /*rerank=*/false, ix->w_ids.p, nullptr, nullptr, nullptr, nullptr, s);
  AB_CUDA(cudaMemcpyAsync(res.data(), ix->w_ids.p, 32 * 4, cudaMemcpyDeviceToHost, s));
  AB_CUDA(cudaStreamSynchronize(s));

  // 2. Graph::insert / space insert: full -> (IDType)-1 (sequential_storage.hpp:77-80)
  if (ix->n >= ix->capacity) return kEmptyId;
  const uint32_t id = static_cast<uint32_t>(ix->n);
  AB_CUDA(cudaMemcpyAsync(ix->adj.p + size_t(id) * 32, res.data(), 32 * 4, cudaMemcpyHostToDevice, s));
  if (raw) {
    stage_rows(ix->w_queries.p, ix->raw.p + size_t(id) * ix->raw_stride, 1, ix->dim, ix->raw_stride,
               ix->params.metric == AB_METRIC_COS, s);
    ix->h_valid_raw[id] = 1;
    ix->n_raw = id + 1;
    if (ix->has_removed) index_sync_valid_bitmap(ix);
  } else {
    // SQ8Space::insert (sq8_space.hpp:188-196): encode with the fitted min/max; the raw space is NOT extended
    quant_encode(quant, ix->w_queries.p, 1, ix->dim, ix->dim, ix->qmin.p, ix->qmax.p, ix->codes.p + size_t(id) * ix->code_stride,
                 ix->code_stride, s);
  }
  ix->h_valid_space[id] = 1;
  ix->h_valid_graph[id] = 1;
  ix->h_levels[id] = 0;
  ix->n = id + 1;

  // 3. reverse edges, then re-rank every touched row (graph_update_job.hpp:77-88)
  for (uint32_t r : res)
    if (r != kEmptyId) ix->inserted_edges[r].push_back(id);
  std::vector<uint32_t> nodes;
  for (auto &kv : ix->inserted_edges) nodes.push_back(kv.first);
  const uint32_t m = static_cast<uint32_t>(nodes.size());
  std::vector<uint32_t> rows(size_t(m) * 32);
  for (uint32_t t = 0; t < m; t++)
    AB_CUDA(cudaMemcpyAsync(rows.data() + size_t(t) * 32, ix->adj.p + size_t(nodes[t]) * 32, 128, cudaMemcpyDeviceToHost, s));
  AB_CUDA(cudaStreamSynchronize(s));
  std::vector<uint32_t> offsets(m + 1, 0), cands;
  for (uint32_t t = 0; t < m; t++) {
    std::unordered_set<uint32_t> cset;  // same container, same insertion sequence as the reference
    for (int j = 0; j < 32; j++) {
      const uint32_t nb = rows[size_t(t) * 32 + j];
      if (nb == kEmptyId) break;
      if (ix->removed_vertices.count(nb))
        for (uint32_t sh : ix->removed_node_nbrs.at(nb)) cset.insert(sh);
      cset.insert(nb);
    }
    for (uint32_t x : ix->inserted_edges.at(nodes[t])) cset.insert(x);
    for (uint32_t x : cset) cands.push_back(x);
    offsets[t + 1] = static_cast<uint32_t>(cands.size());
  }
  std::vector<float> dist;
  node_distances(ix, nodes, offsets, cands, dist);
  for (uint32_t t = 0; t < m; t++) {
    // LinearPool(capacity 32): the 32 closest, equal distances in arrival order
    std::vector<std::pair<float, uint32_t>> order;
    for (uint32_t e = offsets[t]; e < offsets[t + 1]; e++) {
      float d = dist[e];
      if (raw && !ix->h_valid_raw[cands[e]]) d = FLT_MAX;  // raw_space.hpp:298-300
      order.emplace_back(d, e);
    }
    std::stable_sort(order.begin(), order.end(), [](const auto &a, const auto &b) { return a.first < b.first; });
    uint32_t upd[32] = {0};  // value-initialised in the reference: unfilled slots are 0, not -1
    for (size_t i = 0; i < order.size() && i < 32; i++) upd[i] = cands[order[i].second];
    if (!ix->h_valid_graph[nodes[t]]) continue;  // SequentialStorage::update skips invalid rows
    AB_CUDA(cudaMemcpyAsync(ix->adj.p + size_t(nodes[t]) * 32, upd, 128, cudaMemcpyHostToDevice, s));
    AB_CUDA(cudaStreamSynchronize(s));
  }
  ix->inserted_edges.clear();
  AB_CUDA(cudaStreamSynchronize(s));
  return id;
}

void index_remove(ab_index *ix, uint32_t id) {
  std::lock_guard<std::mutex> lock(ix->mu);
  AB_CUDA(cudaSetDevice(ix->device));
  AB_CHECK(ix->has_graph, AB_ERR_STATE, "Index is not init yet");
  AB_CHECK(id < ix->capacity, AB_ERR_INVALID, "id out of range");
  // searches enqueued on other streams read adj / valid: let them finish first
  if (ix->ws_event_recorded) AB_CUDA(cudaEventSynchronize(ix->ws_event));
  uint32_t row[32];
  AB_CUDA(cudaMemcpy(row, ix->adj.p + size_t(id) * 32, 128, cudaMemcpyDeviceToHost));
  for (int j = 0; j < 32; j++) {
    if (row[j] == kEmptyId) break;
    ix->removed_node_nbrs[id].push_back(row[j]);
  }
  ix->removed_vertices.insert(id);
  ix->h_valid_graph[id] = 0;  // Graph::remove
  ix->delete_cnt_space++;     // space_->remove(id): delete_cnt_++ then clear the validity bit
  ix->h_valid_space[id] = 0;
  if (ix->params.quantization_type == AB_QUANT_NONE) {
    // the raw space IS the search space: its QueryComputer now returns FLT_MAX for this row
    ix->delete_cnt_raw++;
    ix->h_valid_raw[id] = 0;
    ix->has_removed = true;
    index_sync_valid_bitmap(ix);
  }
}

}  // namespace ab
