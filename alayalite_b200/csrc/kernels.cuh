// Host-side launchers of the stand-alone kernels (quant.cu, bruteforce.cu, merge in bruteforce.cu).
#pragma once
#include <mutex>

#include "common.cuh"

namespace ab {

// quant.cu
void stage_rows(const float *d_src, float *d_dst, uint64_t n, uint32_t dim, uint32_t stride_floats, bool normalize,
                cudaStream_t s);
void quant_fit(const float *d_rows, uint64_t n, uint32_t dim, uint32_t stride_floats, float *d_min, float *d_max,
               cudaStream_t s);
void quant_encode(int quant, const float *d_rows, uint64_t n, uint32_t dim, uint32_t stride_floats, const float *d_min,
                  const float *d_max, uint8_t *d_codes, uint32_t code_stride, cudaStream_t s);
void quant_tables(int quant, const float *d_min, const float *d_max, uint32_t dim, float *d_scale, float *d_vmin,
                  uint32_t table_len, cudaStream_t s);
void distance_pairs(int metric, int quant, const uint8_t *d_x, const uint8_t *d_y, size_t row_stride, int nchunks,
                    uint64_t pairs, const float *d_scale, const float *d_vmin, int table_len, float *d_out,
                    cudaStream_t s);

// bruteforce.cu
// rows: [n, stride_floats] with stride a multiple of 16 floats and zero padding; queries likewise.
void bruteforce_topk(int metric, const float *d_rows, uint64_t n, uint32_t stride_floats, const float *d_queries,
                     uint64_t nq, uint32_t q_stride_floats, uint32_t k, uint32_t *d_out_ids, float *d_out_dist,
                     int sm_count, cudaStream_t s);
// serialises the users of the persistent brute-force workspaces (both paths); results are stream-ordered on `s`,
// the caller synchronises
std::mutex &bruteforce_mutex();
// bruteforce_tc.cu: tcgen05 TF32 filter + exact re-evaluation (same contract and result as bruteforce_topk)
bool bruteforce_tc_applicable(uint64_t n, uint64_t nq, uint32_t k);
void bruteforce_topk_tc(int metric, const float *d_rows, uint64_t n, uint32_t dim, uint32_t stride_floats, const float *d_queries,
                        uint64_t nq, uint32_t q_stride_floats, uint32_t k, uint32_t *d_out_ids, float *d_out_dist, int sm_count,
                        cudaStream_t s);
void merge_topk(const float *d_dist, const uint32_t *d_ids, uint32_t lists, uint64_t nq, uint32_t k, uint32_t *d_out_ids,
                float *d_out_dist, int sm_count, cudaStream_t s);

}  // namespace ab
