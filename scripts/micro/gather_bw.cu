// Micro-benchmark: throughput of random row gathers from HBM as a function of row size and row stride
// (is a 384-byte row -- d = 96 fp32, the C5 shape -- inherently slower to gather than a 512-byte one?).
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o gather_bw gather_bw.cu && ./gather_bw
#include <cstdint>
#include <cstdio>
#include <cuda_runtime.h>

__global__ void __launch_bounds__(256, 3) gather(const uint8_t *rows, size_t stride, int chunks, const uint32_t *ids, uint32_t n_ids,
                                                 int per_warp, float *out) {
  // 8 lanes per row, 4 rows per warp pass, 3 passes in flight (like the traversal kernel)
  const int lane = threadIdx.x & 31, g = lane >> 3, lg = lane & 7;
  const uint32_t warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  float acc = 0.f;
  for (int it = 0; it < per_warp; it += 12) {
    uint4 x[3][4];
#pragma unroll
    for (int c = 0; c < 3; c++) {
      const uint32_t id = ids[(size_t(warp) * per_warp + it + c * 4 + g) % n_ids];
      const uint8_t *r = rows + size_t(id) * stride;
#pragma unroll
      for (int i = 0; i < 4; i++) {
        if (lg + i * 8 < chunks) {
          asm volatile("ld.global.nc.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(x[c][i].x), "=r"(x[c][i].y), "=r"(x[c][i].z), "=r"(x[c][i].w) : "l"(r + (lg + i * 8) * 16));
        } else {
          x[c][i] = make_uint4(0, 0, 0, 0);
        }
      }
    }
#pragma unroll
    for (int c = 0; c < 3; c++)
#pragma unroll
      for (int i = 0; i < 4; i++) acc += __uint_as_float(x[c][i].x) + __uint_as_float(x[c][i].w);
  }
  if (acc == 12345.678f) out[0] = acc;
}

int main() {
  const size_t n_rows = 12500000;
  uint8_t *rows;
  cudaMalloc(&rows, n_rows * 512);
  cudaMemset(rows, 0, n_rows * 512);
  const uint32_t n_ids = 1u << 24;
  uint32_t *h = new uint32_t[n_ids], *ids;
  uint64_t s = 88172645463325252ull;
  for (uint32_t i = 0; i < n_ids; i++) {
    s ^= s << 13; s ^= s >> 7; s ^= s << 17;
    h[i] = uint32_t(s % n_rows);
  }
  cudaMalloc(&ids, n_ids * 4);
  cudaMemcpy(ids, h, n_ids * 4, cudaMemcpyHostToDevice);
  float *out;
  cudaMalloc(&out, 4);
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0);
  cudaEventCreate(&e1);
  const int per_warp = 1200, blocks = 148 * 3;
  struct Case { int row_bytes; size_t stride; } cases[] = {{512, 512}, {384, 384}, {384, 512}, {256, 256}, {128, 128}, {960, 960}, {1024, 1024}};
  for (auto c : cases) {
    const int chunks = c.row_bytes / 16;
    float best = 1e9f;
    for (int rep = 0; rep < 5; rep++) {
      cudaEventRecord(e0);
      gather<<<blocks, 256>>>(rows, c.stride, chunks > 32 ? 32 : chunks, ids, n_ids, per_warp, out);
      cudaEventRecord(e1);
      cudaEventSynchronize(e1);
      float ms;
      cudaEventElapsedTime(&ms, e0, e1);
      if (ms < best) best = ms;
    }
    const double bytes = double(blocks) * 8 * per_warp * (chunks > 32 ? 512 : c.row_bytes);
    printf("row %4d B  stride %4zu B : %.3f ms  %.0f GB/s of row bytes  (%s)\n", c.row_bytes, c.stride, best, bytes / best / 1e6,
           cudaGetErrorString(cudaGetLastError()));
  }
  return 0;
}
