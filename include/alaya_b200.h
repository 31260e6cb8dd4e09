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
 * build the HNSW graph on the GPU. num_threads is accepted for API compatibility and ignored. */
int ab_index_fit(ab_index *index, const float *data, uint64_t n, uint32_t dim,
                 uint32_t ef_construction, uint32_t num_threads);

/* PyIndex::search (index.hpp:236-258): one query, ids[topk]. */
int ab_index_search(ab_index *index, const float *query, uint32_t topk, uint32_t ef, uint32_t *out_ids);

/* PyIndex::batch_search (index.hpp:289-346): queries [nq, dim] -> ids [nq, topk]. SQ spaces are
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
/* rerank != 0 applies the SQ re-rank of batch_search; 0 returns the search-space pool order.   */
int ab_index_batch_search_device(ab_index *index, const float *d_queries, uint64_t nq, uint32_t topk,
                                 uint32_t ef, int32_t rerank, uint32_t *d_out_ids, float *d_out_dist,
                                 void *stream);

/* Multi-GPU (no counterpart in the reference): the same search, whose kernel additionally stores
 * every final (id + id_offset, distance) into slot `slot` of n_dest gather buffers laid out [slots, nq, topk]
 * -- dest_ids / dest_dist are HOST arrays of n_dest (<= 16) device addresses, normally one buffer per rank of
 * the box mapped into this process as peer memory -- so the shard exchange rides on the search itself and
 * only a barrier + ab_merge_topk_device remain. d_out_ids / d_out_dist receive the local lists as usual. */
int ab_index_batch_search_scatter_device(ab_index *index, const float *d_queries, uint64_t nq, uint32_t topk,
                                         uint32_t ef, int32_t rerank, uint32_t *d_out_ids, float *d_out_dist,
                                         uint32_t id_offset, uint32_t slot, uint32_t n_dest,
                                         const uint64_t *dest_ids, const uint64_t *dest_dist, void *stream);

/* Parity / interchange hooks: split fit into "spaces" and "graph" so that a graph built elsewhere
 * (the reference on the CPU) can be searched on the GPU and vice versa.                          */
int ab_index_fit_spaces(ab_index *index, const float *data, uint64_t n, uint32_t dim);
/* adj [n, R] with 0xFFFFFFFF padding; levels [n]; lists_flat = for every node with level>0, in id
 * order, level*R ids (rows of OverlayGraph::lists_, overlay_graph.hpp:43-44); entry point. */
int ab_index_set_graph(ab_index *index, const uint32_t *adj, const uint32_t *levels,
                       const uint32_t *lists_flat, uint64_t lists_len, uint32_t ep);
int ab_index_build_graph(ab_index *index, uint32_t ef_construction);
/* info = {n, R, ep, lists_len, capacity} */
int ab_index_graph_info(ab_index *index, uint64_t info[5]);
int ab_index_get_graph(ab_index *index, uint32_t *adj, uint32_t *levels, uint32_t *lists_flat);
/* SQ codes [n, row_bytes] (row_bytes = dim for SQ8, (dim+1)/2 for SQ4) and quantizer min/max [dim] */
int ab_index_get_codes(ab_index *index, uint8_t *codes, float *min, float *max);
/* Counters of the last batch search: {level-0 distance evaluations, expanded nodes (hops),
 * overlay distance evaluations, queries re-run through the exact global-bitmap visited set,
 * queries} -- feeds the algorithmic-bytes figure of the roofline. */
int ab_index_search_stats(ab_index *index, uint64_t stats[5]);
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
