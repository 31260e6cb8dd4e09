// Space construction kernels: row staging, cosine normalisation, SQ8 / SQ4 quantizer fit + encode,
// dequantisation tables, and the stand-alone pairwise distance kernel.
//
// Replaces (paths relative to /root/reference):
//   RawSpace::fit + normalize        include/space/raw_space.hpp:131-139, include/utils/data_utils.hpp:36-46
//   SQ8Quantizer / SQ4Quantizer      include/space/quant/sq8.hpp:99-143, include/space/quant/sq4.hpp:99-149
//   SQ8Space::fit / SQ4Space::fit    include/space/sq8_space.hpp:116-127, include/space/sq4_space.hpp:119-130
#include <cfloat>

#include "distance.cuh"
#include "kernels.cuh"

namespace ab {

namespace {

// dense [n, dim] -> padded rows (stride in floats), zero padding; optional L2 normalisation.
// One warp per row. Sum of squares is accumulated per lane then butterfly-reduced (the reference
// accumulates sequentially in one float; the difference is within 1 ulp of the scale factor).
__global__ void stage_rows_kernel(const float *__restrict__ src, float *__restrict__ dst, uint64_t n, uint32_t dim,
                                  uint32_t stride, int normalize) {
  const int lane = threadIdx.x & 31;
  const uint64_t warp = (uint64_t(blockIdx.x) * blockDim.x + threadIdx.x) >> 5;
  const uint64_t nwarps = (uint64_t(gridDim.x) * blockDim.x) >> 5;
  for (uint64_t r = warp; r < n; r += nwarps) {
    const float *s = src + r * dim;
    float *d = dst + r * stride;
    float inv = 1.f;
    if (normalize) {
      float ss = 0.f;
      for (uint32_t i = lane; i < dim; i += 32) {
        const float x = s[i];
        ss = fmaf(x, x, ss);
      }
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) ss += __shfl_xor_sync(0xFFFFFFFFu, ss, o);
      inv = static_cast<float>(1.0 / static_cast<double>(sqrtf(ss)));
    }
    for (uint32_t i = lane; i < stride; i += 32) d[i] = i < dim ? (normalize ? s[i] * inv : s[i]) : 0.f;
  }
}

// per-dimension min / max over rows: pass 1 writes one partial per block, pass 2 folds them.
__global__ void minmax_partial_kernel(const float *__restrict__ rows, uint64_t n, uint32_t dim, uint32_t stride,
                                      float *__restrict__ pmin, float *__restrict__ pmax) {
  for (uint32_t d = threadIdx.x; d < dim; d += blockDim.x) {
    float mn = FLT_MAX, mx = -FLT_MAX;  // sq8.hpp:78-92 initial values
    for (uint64_t r = blockIdx.x; r < n; r += gridDim.x) {
      const float v = rows[r * stride + d];
      mn = fminf(mn, v);
      mx = fmaxf(mx, v);
    }
    pmin[size_t(blockIdx.x) * dim + d] = mn;
    pmax[size_t(blockIdx.x) * dim + d] = mx;
  }
}
__global__ void minmax_final_kernel(const float *__restrict__ pmin, const float *__restrict__ pmax, uint32_t parts,
                                    uint32_t dim, float *__restrict__ mn_out, float *__restrict__ mx_out) {
  const uint32_t d = blockIdx.x * blockDim.x + threadIdx.x;
  if (d >= dim) return;
  float mn = FLT_MAX, mx = -FLT_MAX;
  for (uint32_t p = 0; p < parts; p++) {
    mn = fminf(mn, pmin[size_t(p) * dim + d]);
    mx = fmaxf(mx, pmax[size_t(p) * dim + d]);
  }
  mn_out[d] = mn;
  mx_out[d] = mx;
}

// one thread per output byte; rows of `code_stride` bytes, zero padded
template <int QUANT>
__global__ void encode_kernel(const float *__restrict__ rows, uint64_t n, uint32_t dim, uint32_t stride,
                              const float *__restrict__ mn, const float *__restrict__ mx, uint8_t *__restrict__ codes,
                              uint32_t code_stride) {
  const uint64_t total = n * code_stride;
  for (uint64_t t = uint64_t(blockIdx.x) * blockDim.x + threadIdx.x; t < total; t += uint64_t(gridDim.x) * blockDim.x) {
    const uint64_t r = t / code_stride;
    const uint32_t b = static_cast<uint32_t>(t % code_stride);
    const float *row = rows + r * stride;
    uint32_t out = 0;
    if (QUANT == AB_QUANT_SQ8) {
      if (b < dim) out = sq_quantize(row[b], mn[b], mx[b], 255.f);
    } else {
      const uint32_t i = 2 * b;  // sq4.hpp:139-149: dim i -> high nibble, dim i+1 -> low nibble
      uint32_t first = 0, second = 0;
      if (i < dim) first = sq_quantize(row[i], mn[i], mx[i], 15.f);
      if (i + 1 < dim) second = sq_quantize(row[i + 1], mn[i + 1], mx[i + 1], 15.f);
      out = (first << 4) | second;
    }
    codes[t] = static_cast<uint8_t>(out);
  }
}

// scale[i] = (max[i] - min[i]) * (1 / levels) exactly as the host kernels compute it per call
// (distance_l2.ipp:193-196); zero beyond dim.
__global__ void tables_kernel(const float *__restrict__ mn, const float *__restrict__ mx, uint32_t dim, float inv_levels,
                              float *__restrict__ scale, float *__restrict__ vmin, uint32_t table_len) {
  const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= table_len) return;
  scale[i] = i < dim ? __fmul_rn(__fsub_rn(mx[i], mn[i]), inv_levels) : 0.f;
  vmin[i] = i < dim ? mn[i] : 0.f;
}

// one warp per pair: x staged in shared memory as the "query", y streamed as the candidate row
template <int QUANT, bool IP>
__global__ void pairs_kernel(const uint8_t *__restrict__ x, const uint8_t *__restrict__ y, size_t row_stride, int nchunks,
                             uint64_t pairs, const float *__restrict__ scale, const float *__restrict__ vmin,
                             int table_len, float *__restrict__ out) {
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
  QuerySmem qs;
  qs.q_raw = reinterpret_cast<const float *>(q);
  qs.q_code = q;
  qs.scale = s_scale;
  qs.vmin = s_vmin;
  for (uint64_t p = uint64_t(blockIdx.x) * warps + warp; p < pairs; p += uint64_t(gridDim.x) * warps) {
    for (int c = lane; c < nchunks; c += 32)
      reinterpret_cast<uint4 *>(q)[c] = reinterpret_cast<const uint4 *>(x + p * row_stride)[c];
    __syncwarp();
    const float d = row_distance<QUANT, IP, 32, 4>(y + p * row_stride, nchunks, lane, qs, true);
    if (lane == 0) out[p] = d;
    __syncwarp();
  }
}

}  // namespace

void stage_rows(const float *d_src, float *d_dst, uint64_t n, uint32_t dim, uint32_t stride_floats, bool normalize,
                cudaStream_t s) {
  if (n == 0) return;
  const int threads = 256;
  const int blocks = static_cast<int>(std::min<uint64_t>((n * 32 + threads - 1) / threads, 148 * 16));
  stage_rows_kernel<<<blocks, threads, 0, s>>>(d_src, d_dst, n, dim, stride_floats, normalize ? 1 : 0);
  AB_CUDA(cudaGetLastError());
  count_launch();
}

void quant_fit(const float *d_rows, uint64_t n, uint32_t dim, uint32_t stride_floats, float *d_min, float *d_max,
               cudaStream_t s) {
  const uint32_t parts = static_cast<uint32_t>(std::min<uint64_t>(std::max<uint64_t>(n / 64, 1), 148 * 8));
  DevBuf<float> pmin, pmax;
  pmin.reserve(size_t(parts) * dim);
  pmax.reserve(size_t(parts) * dim);
  minmax_partial_kernel<<<parts, 256, 0, s>>>(d_rows, n, dim, stride_floats, pmin.p, pmax.p);
  AB_CUDA(cudaGetLastError());
  minmax_final_kernel<<<(dim + 127) / 128, 128, 0, s>>>(pmin.p, pmax.p, parts, dim, d_min, d_max);
  AB_CUDA(cudaGetLastError());
  count_launch(2);
  AB_CUDA(cudaStreamSynchronize(s));  // partial buffers are freed on return
}

void quant_encode(int quant, const float *d_rows, uint64_t n, uint32_t dim, uint32_t stride_floats, const float *d_min,
                  const float *d_max, uint8_t *d_codes, uint32_t code_stride, cudaStream_t s) {
  if (n == 0) return;
  const uint64_t total = n * code_stride;
  const int blocks = static_cast<int>(std::min<uint64_t>((total + 255) / 256, 148 * 32));
  if (quant == AB_QUANT_SQ8)
    encode_kernel<AB_QUANT_SQ8><<<blocks, 256, 0, s>>>(d_rows, n, dim, stride_floats, d_min, d_max, d_codes, code_stride);
  else
    encode_kernel<AB_QUANT_SQ4><<<blocks, 256, 0, s>>>(d_rows, n, dim, stride_floats, d_min, d_max, d_codes, code_stride);
  AB_CUDA(cudaGetLastError());
  count_launch();
}

void quant_tables(int quant, const float *d_min, const float *d_max, uint32_t dim, float *d_scale, float *d_vmin,
                  uint32_t table_len, cudaStream_t s) {
  const float inv = quant == AB_QUANT_SQ8 ? 1.0f / 255.0f : 1.0f / 15.0f;
  tables_kernel<<<(table_len + 127) / 128, 128, 0, s>>>(d_min, d_max, dim, inv, d_scale, d_vmin, table_len);
  AB_CUDA(cudaGetLastError());
  count_launch();
}

void distance_pairs(int metric, int quant, const uint8_t *d_x, const uint8_t *d_y, size_t row_stride, int nchunks,
                    uint64_t pairs, const float *d_scale, const float *d_vmin, int table_len, float *d_out,
                    cudaStream_t s) {
  if (pairs == 0) return;
  const int warps = 4;
  const size_t smem = size_t(table_len) * 8 + size_t(warps) * nchunks * 16;
  const int blocks = static_cast<int>(std::min<uint64_t>((pairs + warps - 1) / warps, 148 * 8));
  const bool ip = metric != AB_METRIC_L2;
#define AB_PAIRS(Q, IPV)                                                                                  \
  {                                                                                                       \
    auto kern = pairs_kernel<Q, IPV>;                                                                     \
    AB_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem))); \
    kern<<<blocks, warps * 32, smem, s>>>(d_x, d_y, row_stride, nchunks, pairs, d_scale, d_vmin, table_len, d_out); \
  }
  if (quant == AB_QUANT_NONE) {
    if (ip) AB_PAIRS(AB_QUANT_NONE, true) else AB_PAIRS(AB_QUANT_NONE, false)
  } else if (quant == AB_QUANT_SQ8) {
    if (ip) AB_PAIRS(AB_QUANT_SQ8, true) else AB_PAIRS(AB_QUANT_SQ8, false)
  } else {
    if (ip) AB_PAIRS(AB_QUANT_SQ4, true) else AB_PAIRS(AB_QUANT_SQ4, false)
  }
#undef AB_PAIRS
  AB_CUDA(cudaGetLastError());
  count_launch();
}

}  // namespace ab
