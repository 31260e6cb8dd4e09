// Multi-GPU behind the C ABI: one logical index split by contiguous id ranges over several GPUs (x query replicas).
//
// The reference has no multi-device mode; this is the one addition, and it lives in the library so
// that the reference's single-process pybind host can use every GPU of a box without torch or NCCL:
//
//   ab_comm   one shard of one replica group, owned by the calling process (one process per GPU: torchrun / MPI ...).
//             Every rank allocates one "mailbox" region in its own HBM, publishes it as an opaque handle (CUDA IPC
//             handle + pid; the caller moves the bytes with whatever transport it has), and maps its peers' regions.
//             ab_comm_batch_search_device then runs the local search whose epilogue delivers the lists over NVLink and
//             merges them on the fly (exchange.cuh), plus a one-thread wait for the rows this rank owns.
//   ab_group  a whole S x R grid driven by ONE process (threads only for the parallel build): peers are plain
//             device pointers with peer access enabled, merged rows are written straight into (pinned) host memory by
//             whichever GPU completes a query, and the host just synchronises the streams.
//
// SQ spaces: every shard quantises with the GLOBAL per-dimension min / max (ab_quant_minmax on each shard's rows,
// combined by the caller or by ab_group_fit), so sharded codes equal the codes of the single index.
#include <unistd.h>

#include <algorithm>
#include <cstring>
#include <functional>
#include <memory>
#include <thread>
#include <vector>

#include "index.cuh"
#include "kernels.cuh"

namespace ab {
int guarded_call(const std::function<void()> &f);  // index.cu: exception -> status + ab_last_error
void fit_spaces(ab_index *ix, const float *data, uint64_t n, uint32_t dim, const float *gmin, const float *gmax);
}  // namespace ab

using namespace ab;  // NOLINT

namespace {

constexpr uint32_t kHandleMagic = 0xAB200C01u;

struct HandleBlob {
  uint32_t magic;
  int32_t pid;
  int32_t device;
  uint32_t rank, world;
  uint32_t rows, max_topk;
  uint64_t region_bytes;
  uint64_t ptr;  // address in the owner's process
  cudaIpcMemHandle_t ipc;
};
static_assert(sizeof(HandleBlob) <= AB_COMM_HANDLE_BYTES, "handle blob too large");

// One block behind the search on the rank's stream: thread 0 polls the counter the merging warps of every shard bump
// (system-scope atomics into this rank's memory), then all threads copy the owned merged rows to the caller's buffers
// (device or pinned host memory) -- one launch instead of a wait kernel and two copies.
__global__ void __launch_bounds__(1024) wait_copy_kernel(unsigned long long *done, unsigned long long expect, uint32_t *err,
                                                         const uint32_t *src_ids, const float *src_dist, uint32_t *dst_ids,
                                                         float *dst_dist, size_t words) {
  if (threadIdx.x == 0) {
    const long long t0 = clock64();
    volatile unsigned long long *d = done;
    while (*d < expect) {
      __nanosleep(100);
      if (clock64() - t0 > 6000000000ll) {  // ~3 s: a peer never delivered; report instead of hanging the device
        atomicAdd(err, 1u);
        break;
      }
    }
    *d = 0ull;  // this parity is not touched again before every shard has finished the next call
    __threadfence_system();
  }
  __syncthreads();
  // rows arrive from other GPUs: read at L2. One block is plenty if every thread keeps several 128-bit loads in flight
  // (a word-by-word loop made this copy the largest part of the exchange: ~80 us for 400 KB).
  const bool vec = (reinterpret_cast<uintptr_t>(dst_ids) & 15) == 0 && (dst_dist == nullptr || (reinterpret_cast<uintptr_t>(dst_dist) & 15) == 0);
  size_t done_words = 0;
  if (vec) {
    const size_t quads = words / 4;
    const uint4 *si = reinterpret_cast<const uint4 *>(src_ids), *sd = reinterpret_cast<const uint4 *>(src_dist);
    uint4 *di = reinterpret_cast<uint4 *>(dst_ids), *dd = reinterpret_cast<uint4 *>(dst_dist);
    for (size_t i = threadIdx.x; i < quads; i += size_t(blockDim.x) * 4) {
      uint4 a[4], b[4];
#pragma unroll
      for (int u = 0; u < 4; u++) {
        const size_t j = i + size_t(u) * blockDim.x;
        if (j < quads) {
          a[u] = __ldcg(si + j);
          if (dd) b[u] = __ldcg(sd + j);
        }
      }
#pragma unroll
      for (int u = 0; u < 4; u++) {
        const size_t j = i + size_t(u) * blockDim.x;
        if (j < quads) {
          di[j] = a[u];
          if (dd) dd[j] = b[u];
        }
      }
    }
    done_words = quads * 4;
  }
  for (size_t i = done_words + threadIdx.x; i < words; i += blockDim.x) {
    dst_ids[i] = __ldcg(src_ids + i);
    if (dst_dist) dst_dist[i] = __ldcg(src_dist + i);
  }
}

}  // namespace

struct ab_comm {
  ab_index *ix = nullptr;
  uint32_t rank = 0, world = 1, id_offset = 0;
  uint64_t max_nq = 0;
  uint32_t max_topk = 0, rows = 0;
  int device = 0;
  uint8_t *region = nullptr;
  size_t region_bytes = 0, parity_bytes = 0;
  size_t off_lists = 0, off_arrive = 0, off_done = 0, off_ids = 0, off_dist = 0;
  uint8_t *peer[kMaxShards] = {};
  bool ipc_opened[kMaxShards] = {};
  bool connected = false;
  uint64_t step = 0;
  DevBuf<uint32_t> loc_ids, err;
  DevBuf<float> loc_dist;

  ~ab_comm() {
    cudaSetDevice(device);
    for (uint32_t r = 0; r < world; r++)
      if (ipc_opened[r]) cudaIpcCloseMemHandle(peer[r]);
    if (region) cudaFree(region);
  }
};

namespace {

void comm_layout(ab_comm *c) {
  c->rows = static_cast<uint32_t>((c->max_nq + c->world - 1) / c->world);
  size_t o = 0;
  c->off_lists = o;
  o += round_up(size_t(c->world) * c->rows * c->max_topk * 8, 256);
  c->off_arrive = o;
  o += round_up(size_t(c->rows) * 4, 256);
  c->off_done = o;
  o += 256;
  c->off_ids = o;
  o += round_up(size_t(c->rows) * c->max_topk * 4, 256);
  c->off_dist = o;
  o += round_up(size_t(c->rows) * c->max_topk * 4, 256);
  c->parity_bytes = o;
  c->region_bytes = 2 * o;
}

// Exchange block of one call. out_ids / out_dist == nullptr: merged rows go to the owners' mailbox rows (compact);
// otherwise every owner writes row qi of the given buffers (memory all shards can reach, e.g. mapped host memory).
Exchange comm_exchange(ab_comm *c, uint32_t *out_ids, float *out_dist) {
  Exchange x{};
  x.n_shards = static_cast<int>(c->world);
  x.shard = c->rank;
  x.id_offset = c->id_offset;
  x.rows = c->rows;
  x.compact = out_ids == nullptr ? 1u : 0u;
  const size_t pb = (c->step & 1) * c->parity_bytes;
  for (uint32_t o = 0; o < c->world; o++) {
    uint8_t *b = c->peer[o] + pb;
    x.lists[o] = reinterpret_cast<uint2 *>(b + c->off_lists);
    x.arrive[o] = reinterpret_cast<uint32_t *>(b + c->off_arrive);
    x.done[o] = reinterpret_cast<unsigned long long *>(b + c->off_done);
    x.out_ids[o] = out_ids ? out_ids : reinterpret_cast<uint32_t *>(b + c->off_ids);
    x.out_dist[o] = out_ids ? out_dist : reinterpret_cast<float *>(b + c->off_dist);
  }
  return x;
}

void comm_check_call(ab_comm *c, uint64_t nq, uint32_t topk) {
  AB_CHECK(c->connected, AB_ERR_STATE, "ab_comm_connect has not been called");
  AB_CHECK(nq <= c->max_nq && topk <= c->max_topk, AB_ERR_INVALID, "batch larger than the communicator was created for");
}

// the local search + exchange of one shard, asynchronous on `s`
void comm_search(ab_comm *c, const float *d_queries, uint64_t nq, uint32_t topk, uint32_t ef, uint32_t *out_ids, float *out_dist,
                 cudaStream_t s) {
  c->loc_ids.reserve(nq * topk);
  c->loc_dist.reserve(nq * topk);
  const Exchange x = comm_exchange(c, out_ids, out_dist);
  index_search_device(c->ix, d_queries, nq, topk, ef, /*rerank=*/true, c->loc_ids.p, c->loc_dist.p, nullptr, nullptr, nullptr, s, &x);
  c->step++;
}

}  // namespace

// ---- single-process grid ------------------------------------------------------------------------------
struct ab_group {
  ab_params params{};
  uint32_t S = 1, R = 1;
  std::vector<int> devices;                   // [R * S], replica-major
  std::vector<ab_index *> index;              // [R * S]
  std::vector<std::unique_ptr<ab_comm>> comm; // [R * S]
  std::vector<cudaStream_t> stream;
  uint64_t n = 0, shard_rows = 0;
  uint32_t dim = 0;
  uint64_t comm_nq = 0;
  uint32_t comm_topk = 0;
  // pinned, portable, mapped staging for callers with pageable buffers
  float *h_q = nullptr;
  uint32_t *h_ids = nullptr;
  float *h_dist = nullptr;
  size_t h_q_floats = 0, h_res = 0;
  std::mutex mu;

  ~ab_group() {
    comm.clear();
    for (size_t i = 0; i < index.size(); i++) {
      if (stream.size() > i && stream[i]) {
        cudaSetDevice(devices[i]);
        cudaStreamDestroy(stream[i]);
      }
      if (index[i]) ab_index_destroy(index[i]);
    }
    if (h_q) cudaFreeHost(h_q);
    if (h_ids) cudaFreeHost(h_ids);
    if (h_dist) cudaFreeHost(h_dist);
  }
};

namespace {

void *mapped_host_pointer(const void *p) {
  cudaPointerAttributes at{};
  if (p == nullptr || cudaPointerGetAttributes(&at, p) != cudaSuccess) {
    cudaGetLastError();
    return nullptr;
  }
  return at.type == cudaMemoryTypeHost ? at.devicePointer : nullptr;
}

void group_make_comms(ab_group *g, uint64_t nq, uint32_t topk) {
  if (g->S == 1 || (nq <= g->comm_nq && topk <= g->comm_topk)) return;
  g->comm.clear();
  g->comm_nq = std::max<uint64_t>(nq, g->comm_nq);
  g->comm_topk = std::max(topk, g->comm_topk);
  const uint64_t per_replica = (g->comm_nq + g->R - 1) / g->R;
  std::vector<uint8_t> handles(size_t(g->R) * g->S * AB_COMM_HANDLE_BYTES);
  for (uint32_t r = 0; r < g->R; r++)
    for (uint32_t s = 0; s < g->S; s++) {
      ab_comm *c = nullptr;
      const size_t i = size_t(r) * g->S + s;
      const int rc = ab_comm_create(g->index[i], s, g->S, static_cast<uint32_t>(s * g->shard_rows), per_replica, g->comm_topk, &c,
                                    handles.data() + i * AB_COMM_HANDLE_BYTES);
      AB_CHECK(rc == AB_OK, rc, std::string("ab_comm_create: ") + ab_last_error());
      g->comm.emplace_back(c);
    }
  for (uint32_t r = 0; r < g->R; r++)
    for (uint32_t s = 0; s < g->S; s++) {
      const size_t i = size_t(r) * g->S + s;
      const int rc = ab_comm_connect(g->comm[i].get(), handles.data() + size_t(r) * g->S * AB_COMM_HANDLE_BYTES);
      AB_CHECK(rc == AB_OK, rc, std::string("ab_comm_connect: ") + ab_last_error());
    }
}

}  // namespace

extern "C" {

int ab_comm_create(ab_index *local, uint32_t rank, uint32_t world, uint32_t id_offset, uint64_t max_nq, uint32_t max_topk,
                   ab_comm **out, uint8_t *handle) {
  return guarded_call([&] {
    AB_CHECK(local != nullptr && out != nullptr && handle != nullptr, AB_ERR_INVALID, "null argument");
    AB_CHECK(world >= 1 && world <= static_cast<uint32_t>(kMaxShards) && rank < world, AB_ERR_INVALID, "1..16 shards per replica group");
    AB_CHECK(max_nq >= 1 && max_topk >= 1, AB_ERR_INVALID, "max_nq and max_topk must be positive");
    auto c = std::make_unique<ab_comm>();
    c->ix = local;
    c->rank = rank;
    c->world = world;
    c->id_offset = id_offset;
    c->max_nq = max_nq;
    c->max_topk = max_topk;
    c->device = local->device;
    AB_CUDA(cudaSetDevice(c->device));
    comm_layout(c.get());
    AB_CUDA(cudaMalloc(reinterpret_cast<void **>(&c->region), c->region_bytes));
    AB_CUDA(cudaMemset(c->region, 0, c->region_bytes));
    c->err.alloc_zero(1, nullptr);
    AB_CUDA(cudaDeviceSynchronize());
    HandleBlob b{};
    b.magic = kHandleMagic;
    b.pid = static_cast<int32_t>(getpid());
    b.device = c->device;
    b.rank = rank;
    b.world = world;
    b.rows = c->rows;
    b.max_topk = max_topk;
    b.region_bytes = c->region_bytes;
    b.ptr = reinterpret_cast<uint64_t>(c->region);
    AB_CUDA(cudaIpcGetMemHandle(&b.ipc, c->region));
    std::memset(handle, 0, AB_COMM_HANDLE_BYTES);
    std::memcpy(handle, &b, sizeof(b));
    *out = c.release();
  });
}

int ab_comm_connect(ab_comm *comm, const uint8_t *handles) {
  return guarded_call([&] {
    AB_CHECK(comm != nullptr && handles != nullptr, AB_ERR_INVALID, "null argument");
    AB_CHECK(!comm->connected, AB_ERR_STATE, "communicator is already connected");
    AB_CUDA(cudaSetDevice(comm->device));
    for (uint32_t r = 0; r < comm->world; r++) {
      HandleBlob b{};
      std::memcpy(&b, handles + size_t(r) * AB_COMM_HANDLE_BYTES, sizeof(b));
      AB_CHECK(b.magic == kHandleMagic && b.rank == r && b.world == comm->world, AB_ERR_INVALID, "handle " + std::to_string(r) + " does not belong to this group");
      AB_CHECK(b.rows == comm->rows && b.max_topk == comm->max_topk && b.region_bytes == comm->region_bytes, AB_ERR_INVALID,
               "every rank must create its communicator with the same max_nq / max_topk");
      if (r == comm->rank) {
        comm->peer[r] = comm->region;
      } else if (b.pid == static_cast<int32_t>(getpid())) {
        // same process: plain pointers; different devices need peer access
        if (b.device != comm->device) {
          int can = 0;
          AB_CUDA(cudaDeviceCanAccessPeer(&can, comm->device, b.device));
          AB_CHECK(can, AB_ERR_CUDA, "no peer access between devices " + std::to_string(comm->device) + " and " + std::to_string(b.device));
          const cudaError_t e = cudaDeviceEnablePeerAccess(b.device, 0);
          if (e != cudaSuccess && e != cudaErrorPeerAccessAlreadyEnabled) AB_CUDA(e);
          cudaGetLastError();
        }
        comm->peer[r] = reinterpret_cast<uint8_t *>(b.ptr);
      } else {
        void *p = nullptr;
        AB_CUDA(cudaIpcOpenMemHandle(&p, b.ipc, cudaIpcMemLazyEnablePeerAccess));
        comm->peer[r] = static_cast<uint8_t *>(p);
        comm->ipc_opened[r] = true;
      }
    }
    comm->connected = true;
  });
}

int ab_comm_batch_search_device(ab_comm *comm, const float *d_queries, uint64_t nq, uint32_t topk, uint32_t ef,
                                uint32_t *d_out_ids, float *d_out_dist, void *stream) {
  return guarded_call([&] {
    AB_CHECK(comm != nullptr && d_queries != nullptr && d_out_ids != nullptr, AB_ERR_INVALID, "null argument");
    if (nq == 0) return;
    comm_check_call(comm, nq, topk);
    std::lock_guard<std::mutex> lock(comm->ix->mu);
    AB_CUDA(cudaSetDevice(comm->device));
    cudaStream_t s = static_cast<cudaStream_t>(stream);
    const size_t pb = (comm->step & 1) * comm->parity_bytes;
    comm_search(comm, d_queries, nq, topk, ef, nullptr, nullptr, s);
    // rows this rank owns: queries rank, rank + world, ...
    const uint64_t owned = nq > comm->rank ? (nq - comm->rank + comm->world - 1) / comm->world : 0;
    uint8_t *b = comm->region + pb;
    wait_copy_kernel<<<1, 1024, 0, s>>>(reinterpret_cast<unsigned long long *>(b + comm->off_done), owned, comm->err.p,
                                        reinterpret_cast<const uint32_t *>(b + comm->off_ids), reinterpret_cast<const float *>(b + comm->off_dist),
                                        d_out_ids, d_out_dist, size_t(owned) * topk);
    AB_CUDA(cudaGetLastError());
    count_launch();
  });
}

int ab_comm_errors(ab_comm *comm, uint32_t *out) {
  return guarded_call([&] {
    AB_CHECK(comm != nullptr && out != nullptr, AB_ERR_INVALID, "null argument");
    AB_CUDA(cudaSetDevice(comm->device));
    AB_CUDA(cudaMemcpy(out, comm->err.p, 4, cudaMemcpyDeviceToHost));
  });
}

void ab_comm_destroy(ab_comm *comm) {
  if (comm == nullptr) return;
  cudaSetDevice(comm->device);
  cudaDeviceSynchronize();
  delete comm;
}

// ---- ab_group -------------------------------------------------------------------------------------------
int ab_group_create(const ab_params *params, const int32_t *devices, uint32_t n_shards, uint32_t n_replicas, ab_group **out) {
  return guarded_call([&] {
    AB_CHECK(params != nullptr && devices != nullptr && out != nullptr, AB_ERR_INVALID, "null argument");
    AB_CHECK(n_shards >= 1 && n_shards <= static_cast<uint32_t>(kMaxShards) && n_replicas >= 1 && n_replicas <= 64, AB_ERR_INVALID,
             "1..16 shards x 1..64 replicas");
    auto g = std::make_unique<ab_group>();
    g->params = *params;
    g->S = n_shards;
    g->R = n_replicas;
    const size_t total = size_t(n_shards) * n_replicas;
    g->devices.assign(devices, devices + total);
    g->index.assign(total, nullptr);
    g->stream.assign(total, nullptr);
    for (size_t i = 0; i < total; i++) {
      ab_params p = *params;
      p.device = g->devices[i];
      p.capacity = static_cast<uint32_t>((uint64_t(params->capacity) + n_shards - 1) / n_shards);
      const int rc = ab_index_create(&p, &g->index[i]);
      AB_CHECK(rc == AB_OK, rc, std::string("ab_index_create: ") + ab_last_error());
      AB_CUDA(cudaSetDevice(g->devices[i]));
      AB_CUDA(cudaStreamCreateWithFlags(&g->stream[i], cudaStreamNonBlocking));
    }
    *out = g.release();
  });
}

void ab_group_destroy(ab_group *group) { delete group; }

int ab_group_index(ab_group *group, uint32_t shard, uint32_t replica, ab_index **out) {
  return guarded_call([&] {
    AB_CHECK(group != nullptr && out != nullptr && shard < group->S && replica < group->R, AB_ERR_INVALID, "bad shard / replica");
    *out = group->index[size_t(replica) * group->S + shard];
  });
}

int ab_group_fit(ab_group *group, const float *data, uint64_t n, uint32_t dim, uint32_t ef_construction) {
  return guarded_call([&] {
    AB_CHECK(group != nullptr && data != nullptr && n > 0 && dim > 0, AB_ERR_INVALID, "fit: empty input");
    std::lock_guard<std::mutex> lock(group->mu);
    ab_group *g = group;
    AB_CHECK(n >= g->S, AB_ERR_INVALID, "fewer rows than shards");
    g->n = n;
    g->dim = dim;
    g->shard_rows = (n + g->S - 1) / g->S;
    AB_CHECK(g->shard_rows <= g->index[0]->capacity, AB_ERR_CAPACITY, "The number of data points exceeds the capacity of the space");
    // global quantiser range: per-shard min / max combined on the host
    std::vector<float> gmin, gmax;
    if (g->params.quantization_type != AB_QUANT_NONE) {
      gmin.assign(dim, FLT_MAX);
      gmax.assign(dim, -FLT_MAX);
      std::vector<float> mn(dim), mx(dim);
      for (uint32_t s = 0; s < g->S; s++) {
        const uint64_t r0 = s * g->shard_rows, r1 = std::min(n, r0 + g->shard_rows);
        AB_CUDA(cudaSetDevice(g->devices[s]));
        const int rc = ab_quant_minmax(g->params.metric, data + r0 * dim, r1 - r0, dim, mn.data(), mx.data());
        AB_CHECK(rc == AB_OK, rc, std::string("ab_quant_minmax: ") + ab_last_error());
        for (uint32_t d = 0; d < dim; d++) {
          gmin[d] = std::min(gmin[d], mn[d]);
          gmax[d] = std::max(gmax[d], mx[d]);
        }
      }
    }
    // every (shard, replica) builds its own graph, all GPUs at once
    const size_t total = g->index.size();
    std::vector<int> rcs(total, AB_OK);
    std::vector<std::string> errs(total);
    std::vector<std::thread> th;
    for (size_t i = 0; i < total; i++) {
      th.emplace_back([&, i] {
        const uint32_t s = static_cast<uint32_t>(i % g->S);
        const uint64_t r0 = s * g->shard_rows, r1 = std::min(n, r0 + g->shard_rows);
        rcs[i] = ab_index_fit_ranged(g->index[i], data + r0 * dim, r1 - r0, dim, ef_construction, gmin.empty() ? nullptr : gmin.data(),
                                     gmax.empty() ? nullptr : gmax.data());
        if (rcs[i] != AB_OK) errs[i] = ab_last_error();
      });
    }
    for (auto &t : th) t.join();
    for (size_t i = 0; i < total; i++) AB_CHECK(rcs[i] == AB_OK, rcs[i], "shard build failed: " + errs[i]);
    g->comm.clear();
    g->comm_nq = 0;
    g->comm_topk = 0;
  });
}

int ab_group_batch_search(ab_group *group, const float *queries, uint64_t nq, uint32_t topk, uint32_t ef, uint32_t *out_ids,
                          float *out_dist) {
  return guarded_call([&] {
    AB_CHECK(group != nullptr && queries != nullptr && out_ids != nullptr, AB_ERR_INVALID, "null argument");
    if (nq == 0) return;
    std::lock_guard<std::mutex> lock(group->mu);
    ab_group *g = group;
    AB_CHECK(g->n > 0, AB_ERR_STATE, "index is not fitted or loaded");
    group_make_comms(g, nq, topk);
    // buffers every GPU can reach: the caller's if they are pinned + mapped, else pinned staging
    const float *q_dev = static_cast<const float *>(mapped_host_pointer(queries));
    if (q_dev == nullptr) {
      if (g->h_q_floats < nq * g->dim) {
        if (g->h_q) AB_CUDA(cudaFreeHost(g->h_q));
        AB_CUDA(cudaHostAlloc(reinterpret_cast<void **>(&g->h_q), nq * g->dim * sizeof(float), cudaHostAllocPortable | cudaHostAllocMapped));
        g->h_q_floats = nq * g->dim;
      }
      std::memcpy(g->h_q, queries, nq * g->dim * sizeof(float));
      q_dev = g->h_q;
    }
    uint32_t *ids_dev = static_cast<uint32_t *>(mapped_host_pointer(out_ids));
    float *dist_dev = out_dist ? static_cast<float *>(mapped_host_pointer(out_dist)) : nullptr;
    const bool stage_ids = ids_dev == nullptr, stage_dist = dist_dev == nullptr;  // distances are always produced
    if (g->h_res < nq * topk && (stage_ids || stage_dist)) {
      if (g->h_ids) AB_CUDA(cudaFreeHost(g->h_ids));
      if (g->h_dist) AB_CUDA(cudaFreeHost(g->h_dist));
      AB_CUDA(cudaHostAlloc(reinterpret_cast<void **>(&g->h_ids), nq * topk * 4, cudaHostAllocPortable | cudaHostAllocMapped));
      AB_CUDA(cudaHostAlloc(reinterpret_cast<void **>(&g->h_dist), nq * topk * 4, cudaHostAllocPortable | cudaHostAllocMapped));
      g->h_res = nq * topk;
    }
    if (stage_ids) ids_dev = g->h_ids;
    if (stage_dist) dist_dev = g->h_dist;
    // replica r answers queries [r * per, (r + 1) * per)
    const uint64_t per = (nq + g->R - 1) / g->R;
    for (uint32_t r = 0; r < g->R; r++) {
      const uint64_t q0 = r * per, q1 = std::min(nq, q0 + per);
      if (q0 >= q1) break;
      for (uint32_t s = 0; s < g->S; s++) {
        const size_t i = size_t(r) * g->S + s;
        AB_CUDA(cudaSetDevice(g->devices[i]));
        std::lock_guard<std::mutex> ilock(g->index[i]->mu);
        if (g->S == 1) {
          index_search_device(g->index[i], q_dev + q0 * g->dim, q1 - q0, topk, ef, true, ids_dev + q0 * topk, dist_dev + q0 * topk, nullptr,
                              nullptr, nullptr, g->stream[i]);
        } else {
          comm_check_call(g->comm[i].get(), q1 - q0, topk);
          comm_search(g->comm[i].get(), q_dev + q0 * g->dim, q1 - q0, topk, ef, ids_dev + q0 * topk, dist_dev + q0 * topk, g->stream[i]);
        }
      }
    }
    // every merge runs inside one of these kernels: once all streams are idle, all rows are final
    for (size_t i = 0; i < g->index.size(); i++) {
      AB_CUDA(cudaSetDevice(g->devices[i]));
      AB_CUDA(cudaStreamSynchronize(g->stream[i]));
    }
    if (stage_ids) std::memcpy(out_ids, g->h_ids, nq * topk * 4);
    if (out_dist && stage_dist) std::memcpy(out_dist, g->h_dist, nq * topk * 4);
  });
}

}  // extern "C"
