/*
 * alaya_b200.h -- C ABI of libalaya_b200.so, the B200-native (sm_100a) implementation of
 * AlayaLite's batched k-NN hot path.
 *
 * This is the drop-in boundary: exactly what the reference's pybind layer
 * (/root/reference/python/include/index.hpp, class PyIndex) needs in order to forward
 * fit / search / batch_search / insert / remove / save / load to the GPU. Each entry point names
 * the reference interface it replaces (paths relative to /root/reference). INTEGRATION.md shows
 * the patch a reference maintainer would apply.
 *
 * Conventions
 *   - plain C types only; host pointers in, host pointers out unless the name ends in _device
 *   - every function returns 0 on success or a negative ab_status; ab_last_error() gives the text
 *     (the reference throws std::runtime_error -> Python RuntimeError; the pybind shell re-throws)
 *   - no CPU fallback anywhere: without a CUDA device every compute entry point fails with
 *     AB_ERR_CUDA
 *   - vectors are float32, ids are uint32 (31 usable bits, as in the reference:
 *     include/utils/query_utils.hpp:302-306); other dtypes are converted by the binding layer the
 *     same way the reference's scalar kernels convert them (include/simd/distance_l2.ipp:729-742)
 *   - concurrent searches on one handle are serialised internally; fit/insert/remove/load are not
 *     thread-safe against anything else (same as the reference, SURVEY.md 8b "Threading")
 */
#ifndef ALAYA_B200_H_
#define ALAYA_B200_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct ab_index ab_index;

/* include/utils/metric_type.hpp:27-32 */
enum ab_metric { AB_METRIC_L2 = 0, AB_METRIC_IP = 1, AB_METRIC_COS = 2 };
/* include/utils/quantization_type.hpp */
enum ab_quant { AB_QUANT_NONE = 0, AB_QUANT_SQ8 = 1, AB_QUANT_SQ4 = 2 };
/* include/index/index_type.hpp:27-33. Like the reference's Python path (python/include/index.hpp:209-212)
 * every type builds an HNSW graph; NSG/FUSION only change the saved file name. */
enum ab_index_type { AB_INDEX_FLAT = 0, AB_INDEX_HNSW = 1, AB_INDEX_NSG = 2, AB_INDEX_FUSION = 3 };

enum ab_status {
  AB_OK = 0,
  AB_ERR_INVALID = -1,     /* bad argument (the reference throws std::runtime_error) */
  AB_ERR_CUDA = -2,        /* CUDA runtime error or no device */
  AB_ERR_IO = -3,          /* "Cannot open file ..." (graph.hpp:169-171, raw_space.hpp:221-223) */
  AB_ERR_STATE = -4,       /* index not fitted / loaded */
  AB_ERR_CAPACITY = -5,    /* more rows than capacity (sq8_space.hpp:117-119) */
  AB_ERR_UNSUPPORTED = -6  /* dispatch.hpp:46,90,120,139 "Unsupported ..." */
};

/* python/include/params.hpp:29-52 (IndexParams). max_nbrs is fixed at 32 in the reference's
 * Python path (never forwarded, schema.py:94-101); values other than 32 are rejected. */
typedef struct ab_params {
  int32_t index_type;        /* ab_index_type */
  int32_t quantization_type; /* ab_quant */
  int32_t metric;            /* ab_metric */
  uint32_t capacity;         /* rows the index can ever hold (default 100000 in schema.py:81-82) */
  uint32_t max_nbrs;         /* level-0 degree R; 0 means 32 */
  int32_t device;            /* CUDA device ordinal; -1 = current device */
} ab_params;

/* ------------------------------------------------------------------------------------------- */
/* Index life cycle: PyIndexInterface ctor / dtor (python/include/index.hpp:508-587)             */
int ab_index_create(const ab_params *params, ab_index **out);
void ab_index_destroy(ab_index *index);

/* PyIndex::fit (index.hpp:177-227): copy rows to the device (COS: L2-normalise the device copy;
 * the caller's buffer is NOT modified, unlike raw_space.hpp:131-139), fit + encode the SQ space,
 * build the HNSW graph on the GPU. num_threads is accepted for API compatibility and ignored.
 * `data` may also be device memory on the index's device (rows generated or already resident there). */
int ab_index_fit(ab_index *index, const float *data, uint64_t n, uint32_t dim,
                 uint32_t ef_construction, uint32_t num_threads);

/* PyIndex::search (index.hpp:236-258): one query, ids[topk]. */
int ab_index_search(ab_index *index, const float *query, uint32_t topk, uint32_t ef, uint32_t *out_ids);

/* PyIndex::batch_search (index.hpp:289-346): queries [nq, dim] -> ids [nq, topk] (topk <= ef <= 24576). SQ spaces are
 * re-ranked with exact raw-space distances exactly like index.hpp:337-345 / :450-467. */
int ab_index_batch_search(ab_index *index, const float *queries, uint64_t nq, uint32_t topk,
                          uint32_t ef, uint32_t num_threads, uint32_t *out_ids);

/* PyIndex::batch_search_with_distance (index.hpp:372-418, Linux branch): *has_dist = 1 and
 * out_dist [nq, topk] filled for quantization none; for SQ spaces the un-reranked ids are returned
 * and *has_dist = 0 (the reference returns an empty distance array). */
int ab_index_batch_search_with_distance(ab_index *index, const float *queries, uint64_t nq,
                                        uint32_t topk, uint32_t ef, uint32_t num_threads,
                                        uint32_t *out_ids, float *out_dist, int32_t *has_dist);

/* PyIndex::insert -> GraphUpdateJob::insert_and_update (graph_update_job.hpp:66-89). *out_id is
 * the new row id, or 0xFFFFFFFF when the index is full (index.py:118-123 raises on that). */
int ab_index_insert(ab_index *index, const float *vec, uint32_t ef, uint32_t *out_id);
/* PyIndex::remove -> GraphUpdateJob::remove (graph_update_job.hpp:91-103). */
int ab_index_remove(ab_index *index, uint32_t id);

/* PyIndex::get_data_by_id (index.hpp:98-109): copies dim floats; "id out of range" -> AB_ERR_INVALID. */
int ab_index_get_data_by_id(ab_index *index, uint32_t id, float *out);
/* PyIndexInterface::get_data_dim (index.hpp:582) */
uint32_t ab_index_get_dim(const ab_index *index);
uint64_t ab_index_get_size(const ab_index *index);

/* PyIndex::save / load (index.hpp:113-175): byte-compatible with Graph::save (graph.hpp:165-200),
 * OverlayGraph::save (overlay_graph.hpp:183-194), RawSpace/SQ8Space/SQ4Space::save
 * (raw_space.hpp:235-250, sq8_space.hpp:235-251), SequentialStorage::save
 * (sequential_storage.hpp:110-119). Empty / NULL data_path or quant_path are skipped like there. */
int ab_index_save(ab_index *index, const char *index_path, const char *data_path, const char *quant_path);
int ab_index_load(ab_index *index, const char *index_path, const char *data_path, const char *quant_path);

/* ------------------------------------------------------------------------------------------- */
/* Device-resident entry points (inputs/outputs already in HBM on the index's device).          */
/* d_out_dist may be NULL. `stream` is a cudaStream_t; the call is asynchronous on it            */
/* (NULL = the legacy default stream).                                                         */
/* rerank: bit 0 applies the SQ re-rank of batch_search (0 returns the search-space pool order);  */
/* bit 1 marks list slots the index could not fill as (0xFFFFFFFF, FLT_MAX) instead of the        */
/* reference's zero padding and leaves the (ef - topk) zero slots out of the re-rank -- for       */
/* per-shard lists that are merged afterwards (ab_merge_topk_device skips such entries).          */
int ab_index_batch_search_device(ab_index *index, const float *d_queries, uint64_t nq, uint32_t topk,
                                 uint32_t ef, int32_t rerank, uint32_t *d_out_ids, float *d_out_dist,
                                 void *stream);

/* Parity / interchange hooks: split fit into "spaces" and "graph" so that a graph built elsewhere
 * (the reference on the CPU) can be searched on the GPU and vice versa.                          */
int ab_index_fit_spaces(ab_index *index, const float *data, uint64_t n, uint32_t dim);
/* adj [n, R] with 0xFFFFFFFF padding; levels [n]; lists_flat = for every node with level>0, in id
 * order, level*R ids (rows of OverlayGraph::lists_, overlay_graph.hpp:43-44); entry point. */
int ab_index_set_graph(ab_index *index, const uint32_t *adj, const uint32_t *levels,
                       const uint32_t *lists_flat, uint64_t lists_len, uint32_t ep);
int ab_index_build_graph(ab_index *index, uint32_t ef_construction);
/* NSGBuilder::build_graph (include/index/graph/nsg/nsg_builder.hpp:97-161) on the device, after ab_index_fit_spaces:
 * NN-Descent k-NN graph (include/index/graph/knng/nndescent.hpp:142-400; K = 32, S = 10, nndescent_iters iterations,
 * 0 = the reference's 10) -> navigating node = nearest row to the centroid -> per node the candidates of a traversal
 * of the k-NN graph at ef_construction pruned with the MRNG rule to <= 32 edges -> reverse links -> connectivity pass.
 * The result is a flat graph (one entry point, no overlay) in the reference's Graph layout and file format. Only
 * reachable from C++ in the reference (PyIndex::fit builds HNSW for every index type, and so does ab_index_fit). */
int ab_index_build_graph_nsg(ab_index *index, uint32_t ef_construction, uint32_t nndescent_iters);
/* NndescentImpl::build_graph alone: out_ids [n, 32], the k nearest neighbours found for every row in ascending
 * distance, 0xFFFFFFFF beyond k (k <= 32). */
int ab_index_knng_nndescent(ab_index *index, uint32_t k, uint32_t iters, uint32_t *out_ids);
/* info = {n, R, ep, lists_len, capacity} */
int ab_index_graph_info(ab_index *index, uint64_t info[5]);
int ab_index_get_graph(ab_index *index, uint32_t *adj, uint32_t *levels, uint32_t *lists_flat);
/* SQ codes [n, row_bytes] (row_bytes = dim for SQ8, (dim+1)/2 for SQ4) and quantizer min/max [dim] */
int ab_index_get_codes(ab_index *index, uint8_t *codes, float *min, float *max);
/* Counters of the last batch search: {level-0 distance evaluations, expanded nodes (hops),
 * overlay distance evaluations, queries whose visited set continued in the global overflow table,
 * queries} -- feeds the algorithmic-bytes figure of the roofline. */
int ab_index_search_stats(ab_index *index, uint64_t stats[5]);
/* The same counters, extended: stats[0..4] as above, stats[5] = queries that went past the per-warp overflow
 * table into a pooled bitmap, stats[6] = ids stored in overflow tables, stats[7] = bytes of global memory the
 * visited-set workspace holds (it follows ef and the number of resident warps, not the number of rows).
 * Writes min(count, 8) values, zero beyond. */
int ab_index_search_stats_ex(ab_index *index, uint64_t *stats, uint32_t count);
/* Search-space pool of one batch, before re-rank: pool_ids/pool_dist [nq, ef], sizes [nq].      */
int ab_index_batch_traverse(ab_index *index, const float *queries, uint64_t nq, uint32_t ef,
                            uint32_t *pool_ids, float *pool_dist, uint32_t *pool_sizes);

/* ------------------------------------------------------------------------------------------- */
/* Stand-alone kernels                                                                          */
/* simd::l2_sqr / ip_sqr [_sq8|_sq4] (include/simd/distance_l2.ipp:678-803,
 * distance_ip.ipp:687-816) evaluated on the device for `pairs` row pairs: x,y are [pairs, row]
 * float (quant none) or uint8 codes; min/max [dim] for SQ. metric COS == IP. */
int ab_distance_pairs(int32_t metric, int32_t quant, const void *x, const void *y, uint64_t pairs,
                      uint32_t dim, const float *min, const float *max, float *out);
/* SQ8Quantizer / SQ4Quantizer fit + encode (include/space/quant/sq8.hpp:99-143, sq4.hpp:99-149) */
int ab_quant_fit_encode(int32_t quant, const float *data, uint64_t n, uint32_t dim, float *min,
                        float *max, uint8_t *codes);
/* normalize (include/utils/data_utils.hpp:36-46) applied to n rows in place */
int ab_normalize_rows(float *data, uint64_t n, uint32_t dim);

/* find_exact_gt (include/utils/evaluate.hpp:30-62) / calc_gt (python/src/alayalite/utils.py:99-105):
 * exact top-k of every query against every row. Results are ordered by (distance, id); the
 * reference's order among exactly equal distances is unspecified (std::sort / np.argsort). */
int ab_bruteforce_topk(int32_t metric, const float *data, uint64_t n, uint32_t dim,
                       const float *queries, uint64_t nq, uint32_t k, uint32_t *out_ids,
                       float *out_dist);
int ab_bruteforce_topk_device(int32_t metric, const float *d_data, uint64_t n, uint32_t dim,
                              const float *d_queries, uint64_t nq, uint32_t k, uint32_t *d_out_ids,
                              float *d_out_dist, void *stream);

/* Multi-GPU (no counterpart in the reference): merge `lists` per-shard top-k lists
 * laid out [lists, nq, k] (already carrying global ids) into [nq, k] by (distance, id). */
int ab_merge_topk_device(const float *d_dist, const uint32_t *d_ids, uint32_t lists, uint64_t nq,
                         uint32_t k, uint32_t *d_out_ids, float *d_out_dist, void *stream);

/* ------------------------------------------------------------------------------------------- */
/* Multi-GPU (no counterpart in the reference): one logical index split by contiguous id ranges   */
/* over S GPUs ("shards"), optionally R query replicas of that. Shard s owns rows                               */
/* [s * ceil(n/S), min(n, (s+1) * ceil(n/S))) with local ids, every shard answers every query of its replica    */
/* group, and the per-shard top-k lists are folded into one by (distance, global id). The exchange is fused     */
/* into the search: the kernel that produces a query's final list stores it into the mailbox of the query's     */
/* owner shard (peer memory over NVLink) and the last shard to arrive merges -- no collective, no barrier, no   */
/* merge kernel (csrc/exchange.cuh). SQ spaces must be fitted with the GLOBAL quantiser range so that sharded   */
/* codes equal the codes of the single index.                                                                   */

/* Per-dimension min / max of rows the way SQ8Quantizer::fit sees them (sq8.hpp:99-118; rows L2-normalised first
 * for AB_METRIC_COS because the reference fits its SQ space on the normalised buffer, index.hpp:197-206).
 * `data` may be host or device memory. Combine the per-shard results with min / max and pass them to
 * ab_index_fit_ranged. */
int ab_quant_minmax(int32_t metric, const float *data, uint64_t n, uint32_t dim, float *min, float *max);
/* ab_index_fit with a given quantiser range (min / max [dim], both NULL = fit on these rows as usual). */
int ab_index_fit_ranged(ab_index *index, const float *data, uint64_t n, uint32_t dim, uint32_t ef_construction,
                        const float *min, const float *max);

/* (a) One process drives the whole S x R grid (what the reference's single-process pybind host needs).
 * devices: n_shards * n_replicas CUDA ordinals, replica-major (the shards of replica 0 first); an ordinal may
 * repeat. params->capacity is the capacity of the LOGICAL index. fit builds all shard graphs concurrently;
 * batch_search splits the queries over the replicas, ids are global row numbers of `data`. Pinned + mapped
 * caller buffers are read / written in place by the kernels of every GPU. out_dist may be NULL. */
typedef struct ab_group ab_group;
int ab_group_create(const ab_params *params, const int32_t *devices, uint32_t n_shards, uint32_t n_replicas, ab_group **out);
void ab_group_destroy(ab_group *group);
int ab_group_fit(ab_group *group, const float *data, uint64_t n, uint32_t dim, uint32_t ef_construction);
int ab_group_batch_search(ab_group *group, const float *queries, uint64_t nq, uint32_t topk, uint32_t ef,
                          uint32_t *out_ids, float *out_dist);
/* the ab_index of one shard (owned by the group) */
int ab_group_index(ab_group *group, uint32_t shard, uint32_t replica, ab_index **out);

/* (b) One process per GPU (torchrun, MPI ...): every rank of a replica group creates a communicator around its
 * fitted shard and publishes `handle` (AB_COMM_HANDLE_BYTES opaque bytes: a CUDA IPC handle of the rank's mailbox);
 * the caller gathers the handles of ranks 0..world-1 in rank order with whatever transport it has and every rank
 * connects. max_nq / max_topk bound later calls and must be equal on all ranks.
 * ab_comm_batch_search_device: d_queries [nq, dim] on this rank's device (the same queries on every rank of the
 * group); asynchronous on `stream`. This rank OWNS queries rank, rank + world, ...: their merged lists (global ids
 * = local id + id_offset of the shard that holds the row) arrive in d_out_ids / d_out_dist as
 * [ceil((nq - rank) / world), topk] rows, in that order. Buffers may be device or pinned host memory;
 * d_out_dist may be NULL. Every rank must make the same sequence of calls; two communicators must not share one
 * stream. */
#define AB_COMM_HANDLE_BYTES 128
typedef struct ab_comm ab_comm;
int ab_comm_create(ab_index *local, uint32_t rank, uint32_t world, uint32_t id_offset, uint64_t max_nq,
                   uint32_t max_topk, ab_comm **out, uint8_t *handle);
int ab_comm_connect(ab_comm *comm, const uint8_t *handles);
int ab_comm_batch_search_device(ab_comm *comm, const float *d_queries, uint64_t nq, uint32_t topk, uint32_t ef,
                                uint32_t *d_out_ids, float *d_out_dist, void *stream);
/* number of calls whose wait for the peers timed out (a peer that never searched); 0 = healthy */
int ab_comm_errors(ab_comm *comm, uint32_t *out);
void ab_comm_destroy(ab_comm *comm);

/* Launch plan the traversal would use (host-side decision only, no device needed): for quantization / row
 * length in 16-byte chunks / ef / batch size / rows in the index on a device with `sm_count` SMs and
 * `smem_limit` bytes of opt-in shared memory per block. out = {warps per CTA, resident CTAs per SM, CTAs,
 * visited-table entries per warp, bytes per entry (2 = quotient entries, 4 = ids), dynamic shared memory per CTA}.
 * Exposed so that the occupancy / carve-out policy is testable without a GPU. */
int ab_debug_search_plan(int32_t quant, uint32_t row_chunks, uint32_t ef, uint64_t nq, uint64_t n_rows,
                         int32_t sm_count, uint64_t smem_limit, uint64_t out[6]);

/* Utilities */
const char *ab_last_error(void);
int ab_device_count(void);
int ab_set_device(int32_t device);
/* Number of kernels this library has launched since load (bench.py's gpu_launches). */
uint64_t ab_kernel_launch_count(void);
const char *ab_version(void);

#ifdef __cplusplus
}
#endif
#endif /* ALAYA_B200_H_ */
