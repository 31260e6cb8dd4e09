// Microbenchmark: dependent random 512 B row gathers, the memory pattern of beam_search_kernel.
// Each warp loops: pick 8 pseudo-random rows (seeded by the previous iteration's data), every lane issues
// `INFLIGHT` 128-bit loads (TPV lanes per row), reduce, repeat. Reports ns per dependent round and GB/s.
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
#define CK(x) do{cudaError_t e=(x); if(e!=cudaSuccess){printf("%s: %s\n",#x,cudaGetErrorString(e)); return 1;}}while(0)

template <int MODE> __device__ __forceinline__ uint4 ld(const uint4* p){
  uint4 r;
  if (MODE==0) asm volatile("ld.global.nc.L1::no_allocate.v4.u32 {%0,%1,%2,%3}, [%4];":"=r"(r.x),"=r"(r.y),"=r"(r.z),"=r"(r.w):"l"(p));
  else if (MODE==1) asm volatile("ld.global.v4.u32 {%0,%1,%2,%3}, [%4];":"=r"(r.x),"=r"(r.y),"=r"(r.z),"=r"(r.w):"l"(p));
  else asm volatile("ld.global.nc.v4.u32 {%0,%1,%2,%3}, [%4];":"=r"(r.x),"=r"(r.y),"=r"(r.z),"=r"(r.w):"l"(p));
  return r;
}
// ROWS rows per round, each row 512 B = 32 chunks; TPV lanes per row -> CPL = 32/TPV chunks per lane per row;
// rows per warp-round = (32/TPV) * K where K rows per group in flight.
template <int MODE, int TPV, int K>
__global__ void gather(const uint4* __restrict__ data, uint32_t nrows, int iters, uint32_t* out){
  constexpr int CPL = 32 / TPV, G = 32 / TPV;
  const int lane = threadIdx.x & 31, g = lane / TPV, lg = lane % TPV;
  uint32_t state = (blockIdx.x * blockDim.x + threadIdx.x) / 32 * 2654435761u + 12345u;
  uint32_t acc = 0;
  for (int it = 0; it < iters; it++) {
    uint4 v[K][CPL];
#pragma unroll
    for (int k = 0; k < K; k++) {
      uint32_t h = (state + (g * K + k) * 0x9E3779B9u); h ^= h >> 15; h *= 0x2C1B3C6Du; h ^= h >> 12;
      const uint32_t row = h % nrows;
#pragma unroll
      for (int c = 0; c < CPL; c++) v[k][c] = ld<MODE>(data + size_t(row) * 32 + lg + c * TPV);
    }
    uint32_t s = 0;
#pragma unroll
    for (int k = 0; k < K; k++)
#pragma unroll
      for (int c = 0; c < CPL; c++) s += v[k][c].x ^ v[k][c].y ^ v[k][c].z ^ v[k][c].w;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xFFFFFFFFu, s, o);
    state = state * 1664525u + 1013904223u + s;   // dependent chain: next rows depend on this round's data
    acc += s;
  }
  if (lane == 0) out[(blockIdx.x * blockDim.x + threadIdx.x) / 32] = acc;
}

template <int MODE, int TPV, int K>
int run(const uint4* d, uint32_t nrows, int warps_per_sm, const char* name){
  const int iters = 2000, sms = 148, threads = 256;
  const int blocks = sms * warps_per_sm / (threads / 32);
  uint32_t* out; CK(cudaMalloc(&out, blocks * 8 * 4));
  cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
  gather<MODE, TPV, K><<<blocks, threads>>>(d, nrows, 50, out);
  CK(cudaDeviceSynchronize());
  cudaEventRecord(a);
  gather<MODE, TPV, K><<<blocks, threads>>>(d, nrows, iters, out);
  cudaEventRecord(b); CK(cudaDeviceSynchronize());
  float ms; cudaEventElapsedTime(&ms, a, b);
  const double rows = double(blocks) * 8 * iters * (32 / TPV) * K;
  printf("%-28s rows=%8u (%5.0f MB) warps/SM=%2d tpv=%2d rows/round=%2d : %7.1f ns/round  %7.1f GB/s  %6.2f Grows/s\n", name, nrows, nrows * 512.0 / 1e6,
         warps_per_sm, TPV, (32 / TPV) * K, ms * 1e6 / iters, rows * 512 / (ms * 1e-3) / 1e9, rows / (ms * 1e-3) / 1e9);
  cudaFree(out); return 0;
}

int main(){
  const uint32_t big = 1250000;   // 640 MB
  uint4* d; CK(cudaMalloc(&d, size_t(big) * 512)); CK(cudaMemset(d, 1, size_t(big) * 512));
  for (uint32_t nrows : {60000u, 1250000u}) {
    for (int w : {8, 16, 24, 32, 48, 64}) {
      run<0, 8, 2>(d, nrows, w, "nc.no_allocate tpv8 k2");
    }
    run<1, 8, 2>(d, nrows, 24, "plain ld      tpv8 k2");
    run<2, 8, 2>(d, nrows, 24, "nc            tpv8 k2");
    run<0, 8, 1>(d, nrows, 24, "nc.na tpv8 k1");
    run<0, 8, 4>(d, nrows, 24, "nc.na tpv8 k4");
    run<0, 32, 4>(d, nrows, 24, "nc.na tpv32 k4");
    run<0, 32, 8>(d, nrows, 24, "nc.na tpv32 k8");
    run<0, 4, 1>(d, nrows, 24, "nc.na tpv4 k1");
  }
  return 0;
}
