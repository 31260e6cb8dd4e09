// Brute-force top-k on the 5th-generation tensor cores (tcgen05 + TMEM + TMA), exact results.
//
// Replaces find_exact_gt (/root/reference/include/utils/evaluate.hpp:30-62) and calc_gt
// (python/src/alayalite/utils.py:99-105) for large inputs: this path really is a dense contraction,
// S[q, x] = <Q[q,:], X[x,:]>, 2 * nq * N * d flops.
//
// Tensor cores only see TF32 (10-bit mantissa), which cannot order near neighbours, so the GEMM is
// used as a FILTER whose survivors are re-evaluated in exact fp32:
//   pass A  exact top-k of every query over a sample of the rows (bf_tile_kernel) -> tau_q, an upper
//           bound of the true k-th distance;
//   pass B  (this kernel) TF32 GEMM tile by tile; fused epilogue reads the accumulator from TMEM and
//           keeps every row whose approximate distance ||q||^2 + ||x||^2 - 2<q,x>_tf32 is
//           <= tau_q + eps_q. eps_q bounds the TF32 error (2^-7 ||q|| max||x||, twice the worst case
//           of two truncated operands), so no true top-k row can be dropped: the survivor list is a
//           superset of the exact answer;
//   pass C  exact difference-form distances of the survivors, top-k by (distance, id).
// A query whose survivor list overflows its buffer is re-run on the exact kernel, so the result is
// always the exact fp32 answer (tests compare it with the oracle bit for bit on integer data).
//
// Kernel B layout (one CTA = 128 queries x a slice of the rows, 256 threads):
//   warp 0      TMA producer: cp.async.bulk.tensor 2D loads of the Q tile (128 x 32 fp32) and the X
//               tile (256 x 32 fp32) per k-block into a 4-stage shared-memory ring, 128B swizzle
//   warp 1      TMEM allocation + single-thread tcgen05.mma.cta_group::1.kind::tf32 issue,
//               M = 128, N = 256, K = 8 per instruction; two 256-column accumulators (all 512 TMEM
//               columns) so the epilogue of tile t overlaps the MMAs of tile t + 1
//   warps 4-11  epilogue: tcgen05.ld 32x32b (thread = one query row x half of the columns); the minimum
//               score of 32 columns is tested against the query's threshold first, survivors (rare) are
//               appended to the query's candidate list
// mbarriers: full/empty per stage (TMA <-> MMA), tmem_full/tmem_empty per accumulator (MMA <-> epilogue).
#include <cuda.h>

#include <cfloat>
#include <cmath>
#include <cstdlib>
#include <mutex>
#include <vector>

#include "distance.cuh"
#include "kernels.cuh"
#include "topk.cuh"

namespace ab {

namespace {

constexpr int TC_M = 128;       // queries per CTA (UMMA M, one TMEM lane each)
constexpr int TC_N = 256;       // rows per tile (UMMA N, TMEM columns of one accumulator)
constexpr int TC_KB = 32;       // fp32 elements per k-block = 128 bytes = one swizzle row
constexpr int TC_STAGES = 4;
constexpr int TC_THREADS = 384;   // warp 0 TMA, warp 1 MMA, warps 4-11 epilogue (two per TMEM lane quadrant)
constexpr uint32_t A_BYTES = TC_M * TC_KB * 4;  // 16 KB
constexpr uint32_t B_BYTES = TC_N * TC_KB * 4;  // 32 KB
constexpr uint32_t STAGE_BYTES = A_BYTES + B_BYTES;
constexpr size_t TC_SMEM = 1024 + size_t(TC_STAGES) * STAGE_BYTES + 2 * TC_N * 4 + TC_M * 4 + 256;

// ---- PTX wrappers -------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void *p) { return static_cast<uint32_t>(__cvta_generic_to_shared(p)); }

__device__ __forceinline__ void mbar_init(uint64_t *bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t *bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t *bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t *bar, uint32_t parity) {
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "WAIT_LOOP:\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
      "@p bra WAIT_DONE;\n\t"
      "bra WAIT_LOOP;\n\t"
      "WAIT_DONE:\n\t"
      "}" ::"r"(smem_u32(bar)),
      "r"(parity)
      : "memory");
}
__device__ __forceinline__ void tma_load_2d(void *dst, const CUtensorMap *map, int c0, int c1, uint64_t *bar) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];" ::"r"(smem_u32(dst)),
      "l"(map), "r"(c0), "r"(c1), "r"(smem_u32(bar))
      : "memory");
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_commit(uint64_t *bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}
// D[tmem] (+)= A[smem] * B[smem]^T, TF32 inputs, fp32 accumulate
__device__ __forceinline__ void tc_mma_tf32(uint32_t tmem_d, uint64_t desc_a, uint64_t desc_b, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t"
      "}" ::"r"(tmem_d),
      "l"(desc_a), "l"(desc_b), "r"(idesc), "r"(accumulate)
      : "memory");
}
// K-major operand tile in shared memory, rows of 128 bytes, 128B swizzle (what the TMA wrote):
// start address >> 4 | LBO = 1 (unused for swizzled K-major) | SBO = 1024 B between 8-row groups |
// descriptor version 1 (Blackwell) | layout type 2 = SWIZZLE_128B
__device__ __forceinline__ uint64_t make_smem_desc(const void *tile) {
  uint64_t d = (smem_u32(tile) >> 4) & 0x3FFFu;
  d |= uint64_t(1) << 16;
  d |= uint64_t(1024 >> 4) << 32;
  d |= uint64_t(1) << 46;
  d |= uint64_t(2) << 61;
  return d;
}
// instruction descriptor: D = F32, A = B = TF32, both K-major, N >> 3, M >> 4
constexpr uint32_t kIdescTf32 = (1u << 4) | (2u << 7) | (2u << 10) | (uint32_t(TC_N >> 3) << 17) | (uint32_t(TC_M >> 4) << 24);

__device__ __forceinline__ void tmem_ld32(uint32_t taddr, uint32_t (&v)[32]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
      "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
      : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]), "=r"(v[9]),
        "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]), "=r"(v[16]), "=r"(v[17]), "=r"(v[18]),
        "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]), "=r"(v[24]), "=r"(v[25]), "=r"(v[26]), "=r"(v[27]),
        "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
      : "r"(taddr));
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
}

__device__ __forceinline__ void tmem_ld32_nowait(uint32_t taddr, uint32_t (&v)[32]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
      "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
      : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]), "=r"(v[9]),
        "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]), "=r"(v[16]), "=r"(v[17]), "=r"(v[18]),
        "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]), "=r"(v[24]), "=r"(v[25]), "=r"(v[26]), "=r"(v[27]),
        "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
      : "r"(taddr));
}

struct TcArgs {
  uint64_t n, nq;
  uint32_t kblocks;        // ceil(dim / 32)
  uint32_t tiles_per_split;
  float alpha;             // -2 (L2: ||x||^2 - 2<q,x>) or -1 (IP: -<q,x>)
  const float *xnorm;      // [n] ||x||^2 for L2, zeros for IP
  const float *thr;        // [nq] threshold on alpha * <q,x> + xnorm (i.e. tau + eps - ||q||^2 for L2)
  uint32_t *cand;          // [splits, nq, cap] survivor ids
  uint32_t *counts;        // [splits, nq] survivors seen (may exceed cap -> overflow)
  uint32_t cap;
};

// RESIDENT: the whole 128-query tile (kblocks <= 4, i.e. d <= 128) is loaded once and stays in shared memory;
// only row tiles stream through the ring (2/3 of the L2 -> SM traffic per MMA). Otherwise both operands stream.
template <bool RESIDENT>
__global__ void __launch_bounds__(TC_THREADS, 1)
bf_tc_kernel(const __grid_constant__ CUtensorMap map_q, const __grid_constant__ CUtensorMap map_x, const TcArgs a) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t *smem = reinterpret_cast<uint8_t *>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  // streaming: [stages][A 16 KB | B 32 KB]; resident: [4 x A 16 KB][stages][B 32 KB]. Every tile 1024-byte aligned.
  constexpr uint32_t kStage = RESIDENT ? B_BYTES : STAGE_BYTES;
  constexpr uint32_t kBOff = RESIDENT ? 0u : A_BYTES;
  uint8_t *res_a = smem;
  uint8_t *stage_base = smem + (RESIDENT ? 4 * A_BYTES : 0);
  float *sx = reinterpret_cast<float *>(smem + size_t(TC_STAGES) * STAGE_BYTES);  // [2][TC_N] row terms
  uint32_t *scnt = reinterpret_cast<uint32_t *>(sx + 2 * TC_N);  // [TC_M] survivors per query row
  uint64_t *bars = reinterpret_cast<uint64_t *>(scnt + TC_M);
  uint64_t *full = bars, *empty = bars + TC_STAGES, *tfull = bars + 2 * TC_STAGES, *tempty = bars + 2 * TC_STAGES + 2;
  uint64_t *abar = bars + 2 * TC_STAGES + 4;
  uint32_t *tmem_slot = reinterpret_cast<uint32_t *>(bars + 2 * TC_STAGES + 5);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t q0 = blockIdx.x * TC_M;
  const uint32_t split = blockIdx.y;
  const uint64_t total_tiles = (a.n + TC_N - 1) / TC_N;
  const uint64_t tile_begin = uint64_t(split) * a.tiles_per_split;
  const uint64_t tile_end = min(total_tiles, tile_begin + a.tiles_per_split);
  const uint32_t ntiles = tile_end > tile_begin ? static_cast<uint32_t>(tile_end - tile_begin) : 0u;

  if (threadIdx.x < TC_M) scnt[threadIdx.x] = 0;
  if (threadIdx.x == 0) {
    for (int s = 0; s < TC_STAGES; s++) {
      mbar_init(&full[s], 1);
      mbar_init(&empty[s], 1);
    }
    for (int b = 0; b < 2; b++) {
      mbar_init(&tfull[b], 1);
      mbar_init(&tempty[b], 256);
    }
    mbar_init(abar, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 1) {  // whole warp: allocate all 512 TMEM columns (two 128 x 256 fp32 accumulators)
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(512u) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    // ===== TMA producer =====
    if (lane == 0) {
      if (RESIDENT && ntiles > 0) {
        mbar_expect_tx(abar, a.kblocks * A_BYTES);
        for (uint32_t kb = 0; kb < a.kblocks; kb++)
          tma_load_2d(res_a + size_t(kb) * A_BYTES, &map_q, static_cast<int>(kb * TC_KB), static_cast<int>(q0), abar);
      }
      uint32_t it = 0;
      for (uint32_t t = 0; t < ntiles; t++) {
        const int r0 = static_cast<int>((tile_begin + t) * TC_N);
        for (uint32_t kb = 0; kb < a.kblocks; kb++, it++) {
          const uint32_t s = it % TC_STAGES, ph = (it / TC_STAGES) & 1u;
          mbar_wait(&empty[s], ph ^ 1u);
          uint8_t *sa = stage_base + size_t(s) * kStage;
          mbar_expect_tx(&full[s], kStage);
          if (!RESIDENT) tma_load_2d(sa, &map_q, static_cast<int>(kb * TC_KB), static_cast<int>(q0), &full[s]);
          tma_load_2d(sa + kBOff, &map_x, static_cast<int>(kb * TC_KB), r0, &full[s]);
        }
      }
    }
  } else if (warp == 1) {
    // ===== MMA issuer (one thread) =====
    if (lane == 0) {
      uint32_t it = 0;
      if (RESIDENT && ntiles > 0) mbar_wait(abar, 0);
      for (uint32_t t = 0; t < ntiles; t++) {
        const uint32_t b = t & 1u;
        mbar_wait(&tempty[b], ((t >> 1) & 1u) ^ 1u);
        tc_fence_after();
        const uint32_t tmem_d = tmem_base + b * TC_N;
        for (uint32_t kb = 0; kb < a.kblocks; kb++, it++) {
          const uint32_t s = it % TC_STAGES, ph = (it / TC_STAGES) & 1u;
          mbar_wait(&full[s], ph);
          tc_fence_after();
          const uint8_t *sa = stage_base + size_t(s) * kStage;
          const uint64_t da = make_smem_desc(RESIDENT ? res_a + size_t(kb) * A_BYTES : sa), db = make_smem_desc(sa + kBOff);
#pragma unroll
          for (uint32_t k = 0; k < TC_KB / 8; k++)  // UMMA K = 8 tf32 = 32 bytes = 2 descriptor units
            tc_mma_tf32(tmem_d, da + 2 * k, db + 2 * k, kIdescTf32, (kb | k) != 0 ? 1u : 0u);
          tc_commit(&empty[s]);  // frees the stage when these MMAs retire
        }
        tc_commit(&tfull[b]);  // accumulator b complete
      }
    }
  } else if (warp >= 4) {
    // ===== epilogue: 256 threads; thread e scans columns [half * 128, half * 128 + 128) of query row q0 + (e & 127) =====
    const uint32_t e = threadIdx.x - 128;
    const uint32_t row = e & 127u, half = e >> 7;
    const uint32_t q = q0 + row;
    const bool qvalid = q < a.nq;
    const float thr = qvalid ? a.thr[q] : -INFINITY;
    uint32_t *my = a.cand + (size_t(split) * a.nq + (qvalid ? q : 0)) * a.cap;
    const uint32_t lane_quadrant = (uint32_t(warp & 3) * 32u) << 16;
    // per-row terms (||x||^2 or 0) of the next tile are fetched while the current one is scanned
    float xn_next = ntiles > 0 ? (tile_begin * TC_N + e < a.n ? a.xnorm[tile_begin * TC_N + e] : INFINITY) : 0.f;
    for (uint32_t t = 0; t < ntiles; t++) {
      const uint32_t b = t & 1u;
      const uint64_t r0 = (tile_begin + t) * TC_N;
      float *sxb = sx + b * TC_N;
      sxb[e] = xn_next;  // rows past n carry +inf and can never pass the threshold
      if (t + 1 < ntiles) {
        const uint64_t rn = r0 + TC_N + e;
        xn_next = rn < a.n ? a.xnorm[rn] : INFINITY;
      }
      asm volatile("bar.sync 1, 256;" ::: "memory");
      mbar_wait(&tfull[b], (t >> 1) & 1u);
      tc_fence_after();
#pragma unroll 1
      for (uint32_t c = 0; c < TC_N / 128; c++) {
        const uint32_t col = half * (TC_N / 2) + c * 64;
        uint32_t v[2][32];
        tmem_ld32_nowait(tmem_base + lane_quadrant + b * TC_N + col, v[0]);
        tmem_ld32_nowait(tmem_base + lane_quadrant + b * TC_N + col + 32, v[1]);
        asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
        for (int h = 0; h < 2; h++) {
          const uint32_t cc = col + h * 32;
          // cheap test first: the minimum approximate score per 8 columns. A hit is rare per thread (~1.5 % of
          // the 32-column groups) but common per warp (~40 %), so the extraction must be short: only the 8-column
          // sub-groups whose minimum passes are rescanned, under a warp-uniform outer branch.
          float m4[4] = {INFINITY, INFINITY, INFINITY, INFINITY};
#pragma unroll
          for (int j = 0; j < 32; j += 4) {
            const float4 s4 = *reinterpret_cast<const float4 *>(sxb + cc + j);
            float &m = m4[j >> 3];
            m = fminf(m, fmaf(a.alpha, __uint_as_float(v[h][j + 0]), s4.x));
            m = fminf(m, fmaf(a.alpha, __uint_as_float(v[h][j + 1]), s4.y));
            m = fminf(m, fmaf(a.alpha, __uint_as_float(v[h][j + 2]), s4.z));
            m = fminf(m, fmaf(a.alpha, __uint_as_float(v[h][j + 3]), s4.w));
          }
          const float mall = fminf(fminf(m4[0], m4[1]), fminf(m4[2], m4[3]));
          if (__any_sync(0xFFFFFFFFu, mall <= thr)) {
#pragma unroll
            for (int g = 0; g < 4; g++) {
              if (m4[g] <= thr) {
                // branch-free mask of the passing columns (independent compares), then one append per set bit
                const float4 sa = *reinterpret_cast<const float4 *>(sxb + cc + g * 8);
                const float4 sb = *reinterpret_cast<const float4 *>(sxb + cc + g * 8 + 4);
                const float sx8[8] = {sa.x, sa.y, sa.z, sa.w, sb.x, sb.y, sb.z, sb.w};
                uint32_t pass = 0;
#pragma unroll
                for (int j = 0; j < 8; j++)
                  pass |= (fmaf(a.alpha, __uint_as_float(v[h][g * 8 + j]), sx8[j]) <= thr ? 1u : 0u) << j;
                while (pass) {
                  const int j = __ffs(pass) - 1;
                  pass &= pass - 1;
                  const uint32_t pos = atomicAdd(&scnt[row], 1u);
                  if (pos < a.cap) my[pos] = static_cast<uint32_t>(r0 + cc + g * 8 + j);
                }
              }
            }
          }
        }
      }
      tc_fence_before();
      mbar_arrive(&tempty[b]);
    }
    asm volatile("bar.sync 1, 256;" ::: "memory");
    if (qvalid && half == 0) a.counts[size_t(split) * a.nq + q] = scnt[row];
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512u) : "memory");
  }
}

// ---- helper kernels --------------------------------------------------------------------------------
// ||row||^2 per row (one warp per row) and the global maximum of it
__global__ void row_norms_kernel(const float *__restrict__ rows, uint64_t n, uint32_t stride, float *__restrict__ out,
                                 unsigned int *__restrict__ max_bits) {
  const int lane = threadIdx.x & 31;
  const uint64_t warp = (uint64_t(blockIdx.x) * blockDim.x + threadIdx.x) >> 5;
  const uint64_t nwarps = (uint64_t(gridDim.x) * blockDim.x) >> 5;
  float mx = 0.f;
  for (uint64_t r = warp; r < n; r += nwarps) {
    float s = 0.f;
    for (uint32_t i = lane; i < stride; i += 32) {
      const float x = rows[r * stride + i];
      s = fmaf(x, x, s);
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xFFFFFFFFu, s, o);
    if (lane == 0) out[r] = s;
    mx = fmaxf(mx, s);
  }
  if (lane == 0) atomicMax(max_bits, __float_as_uint(mx));  // non-negative floats order like their bit patterns
}

// thr[q] = tau_q + eps_q (- ||q||^2 for L2), tau_q = k-th exact distance over the sample
__global__ void thresholds_kernel(const float *__restrict__ sample_dist, const uint32_t *__restrict__ sample_ids, uint32_t k,
                                  const float *__restrict__ qnorm, const unsigned int *__restrict__ max_xnorm_bits, int l2,
                                  uint64_t nq, float *__restrict__ thr) {
  const uint64_t q = uint64_t(blockIdx.x) * blockDim.x + threadIdx.x;
  if (q >= nq) return;
  const bool have = sample_ids[q * k + k - 1] != kEmptyId && sample_dist[q * k + k - 1] < FLT_MAX;
  const float tau = have ? sample_dist[q * k + k - 1] : FLT_MAX;
  const float qn = qnorm[q], xn = __uint_as_float(*max_xnorm_bits);
  // TF32 keeps 10 mantissa bits; both operands truncated -> |err(<q,x>)| <= 2^-9 ||q|| ||x||. The score uses
  // 2<q,x> (L2) or <q,x> (IP); 2^-7 leaves a factor >= 2 of head room, plus a little for fp32 accumulation.
  const float eps = 0.0078125f * sqrtf(qn) * sqrtf(xn) + 1e-6f * (qn + xn);
  if (tau >= FLT_MAX) {
    thr[q] = FLT_MAX;
  } else {
    thr[q] = l2 ? (tau + eps - qn) : (tau + eps);
  }
}

// pass C: exact distances of the survivors, top-k by (distance, id); one warp per query
template <bool IP>
__global__ void __launch_bounds__(128) refine_kernel(const float *__restrict__ rows, uint32_t stride, int nchunks,
                                                     const float *__restrict__ queries, uint32_t qstride, uint64_t nq,
                                                     const uint32_t *__restrict__ cand, const uint32_t *__restrict__ counts,
                                                     uint32_t splits, uint32_t cap, uint32_t k, uint32_t need,
                                                     uint32_t *__restrict__ out_ids, float *__restrict__ out_dist,
                                                     uint32_t *__restrict__ redo_flags, uint32_t *__restrict__ redo_count,
                                                     int threshold_only) {
  extern __shared__ __align__(16) uint8_t smem[];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, warps = blockDim.x >> 5;
  const size_t per_warp = round_up16(size_t(nchunks) * 16 + size_t(k) * 8 + 64 * 4);
  uint8_t *wb = smem + size_t(warp) * per_warp;
  float *qs_raw = reinterpret_cast<float *>(wb);
  float *td = reinterpret_cast<float *>(wb + size_t(nchunks) * 16);
  uint32_t *ti = reinterpret_cast<uint32_t *>(td + k);
  for (uint64_t q = uint64_t(blockIdx.x) * warps + warp; q < nq; q += uint64_t(gridDim.x) * warps) {
    for (int c = lane; c < nchunks; c += 32)
      reinterpret_cast<uint4 *>(qs_raw)[c] = reinterpret_cast<const uint4 *>(queries + q * qstride)[c];
    __syncwarp();
    int size = 0;
    bool overflow = false;
    const int nfl = nchunks * 4;
    for (uint32_t sp = 0; sp < splits; sp++) {
      const uint32_t cnt = counts[size_t(sp) * nq + q];
      if (cnt > cap) overflow = true;
      const uint32_t m = min(cnt, cap);
      const uint32_t *lst = cand + (size_t(sp) * nq + q) * cap;
      for (uint32_t base = 0; base < m; base += 32) {  // one survivor per lane
        const uint32_t kk = base + lane;
        const bool act = kk < m;
        const uint32_t id = act ? lst[kk] : 0u;
        // same arithmetic as bf_tile_kernel: one accumulator, dimension order, FMA -> identical distances
        const float4 *xr = reinterpret_cast<const float4 *>(rows + size_t(id) * stride);
        float acc = 0.f;
        if (act) {
          for (int c = 0; c < nfl / 4; c++) {
            const float4 x = xr[c];
            const float4 qq = reinterpret_cast<const float4 *>(qs_raw)[c];
            if (IP) {
              acc = fmaf(qq.x, x.x, acc);
              acc = fmaf(qq.y, x.y, acc);
              acc = fmaf(qq.z, x.z, acc);
              acc = fmaf(qq.w, x.w, acc);
            } else {
              float df = qq.x - x.x;
              acc = fmaf(df, df, acc);
              df = qq.y - x.y;
              acc = fmaf(df, df, acc);
              df = qq.z - x.z;
              acc = fmaf(df, df, acc);
              df = qq.w - x.w;
              acc = fmaf(df, df, acc);
            }
          }
        }
        const float d = IP ? -acc : acc;
        const uint32_t valid = min(32u, m - base);
        for (uint32_t j = 0; j < valid; j++)
          topk_insert(td, ti, size, static_cast<int>(k), __shfl_sync(0xFFFFFFFFu, d, j), __shfl_sync(0xFFFFFFFFu, id, j), lane);
      }
    }
    // threshold_only (the sample level, whose lists only feed the thresholds of the level above): any k real rows
    // bound the true k-th distance from above, so an overflowed list is still usable as it is; a list with fewer
    // than k rows is marked empty (threshold = +inf, the level above keeps everything for that query and repairs it)
    const uint32_t pad_id = threshold_only ? kEmptyId : 0u;
    const float pad_d = threshold_only ? FLT_MAX : 0.f;
    for (uint32_t i = lane; i < k; i += 32) {
      out_ids[q * k + i] = i < static_cast<uint32_t>(size) ? ti[i] : pad_id;
      if (out_dist) out_dist[q * k + i] = i < static_cast<uint32_t>(size) ? td[i] : pad_d;
    }
    if (lane == 0 && !threshold_only) {
      const bool again = overflow || size < static_cast<int>(need);
      redo_flags[q] = again ? 1u : 0u;
      if (again) atomicAdd(redo_count, 1u);
    }
    __syncwarp();
  }
}

__global__ void gather_rows_f32(const float *__restrict__ src, uint32_t stride, const uint32_t *__restrict__ ids, uint32_t m,
                                float *__restrict__ dst) {
  const uint64_t total = uint64_t(m) * stride;
  for (uint64_t t = uint64_t(blockIdx.x) * blockDim.x + threadIdx.x; t < total; t += uint64_t(gridDim.x) * blockDim.x)
    dst[t] = src[uint64_t(ids[t / stride]) * stride + t % stride];
}
__global__ void scatter_results(const uint32_t *__restrict__ ids_in, const float *__restrict__ dist_in, const uint32_t *__restrict__ which,
                                uint32_t m, uint32_t k, uint32_t *__restrict__ out_ids, float *__restrict__ out_dist) {
  const uint64_t t = uint64_t(blockIdx.x) * blockDim.x + threadIdx.x;
  if (t >= uint64_t(m) * k) return;
  const uint64_t dst = uint64_t(which[t / k]) * k + t % k;
  out_ids[dst] = ids_in[t];
  if (out_dist) out_dist[dst] = dist_in[t];
}

using EncodeFn = CUresult (*)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *, const cuuint64_t *,
                              const cuuint32_t *, const cuuint32_t *, CUtensorMapInterleave, CUtensorMapSwizzle,
                              CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

EncodeFn get_encode() {
  static EncodeFn fn = nullptr;
  if (fn == nullptr) {
    void *p = nullptr;
    cudaDriverEntryPointQueryResult qres;
    AB_CUDA(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &qres));
    AB_CHECK(p != nullptr && qres == cudaDriverEntryPointSuccess, AB_ERR_CUDA, "cuTensorMapEncodeTiled is not available");
    fn = reinterpret_cast<EncodeFn>(p);
  }
  return fn;
}

// [rows, stride] fp32, K-major; box = 32 floats (128 B) x box_rows, 128B swizzle, zero fill outside
CUtensorMap make_map(const float *base, uint64_t rows, uint32_t dim, uint32_t stride, uint32_t box_rows) {
  CUtensorMap m;
  const cuuint64_t gdim[2] = {dim, rows};
  const cuuint64_t gstride[1] = {cuuint64_t(stride) * 4};
  const cuuint32_t box[2] = {TC_KB, box_rows};
  const cuuint32_t estr[2] = {1, 1};
  const CUresult r = get_encode()(&m, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, const_cast<float *>(base), gdim, gstride, box, estr,
                                  CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                                  CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  AB_CHECK(r == CUDA_SUCCESS, AB_ERR_CUDA, "cuTensorMapEncodeTiled failed (" + std::to_string(int(r)) + ")");
  return m;
}

}  // namespace

bool bruteforce_tc_applicable(uint64_t n, uint64_t nq, uint32_t k) {
  if (const char *e = std::getenv("AB_BRUTEFORCE_TC")) return std::atoi(e) != 0 && k <= 256;
  return n >= 131072 && nq >= 256 && k <= 128;
}

struct TcWorkspace {
  DevBuf<float> sd, xnorm, qnorm, thr, rd;
  DevBuf<uint32_t> si, maxbits, cand, counts, redo, ri;
};

// rows / queries padded to a multiple of 16 floats (zero padding). Exact result, see the file header.
// level 0 is the caller's problem; level 1 is pass A of level 0 run through the same pipeline on the sample
// (its own pass A is the CUDA-core kernel on a sample of the sample), which takes pass A of 1M x 128 x 10k
// from 5.0 ms of CUDA-core work to ~1.5 ms.
static void tc_run(int metric, const float *d_rows, uint64_t n, uint32_t dim, uint32_t stride, const float *d_queries,
                   uint64_t nq, uint32_t qstride, uint32_t k, uint32_t *d_out_ids, float *d_out_dist, int sm_count,
                   cudaStream_t s, int level) {
  const bool l2 = metric == AB_METRIC_L2;
  AB_CHECK(n >= k && k <= 256, AB_ERR_INVALID, "bruteforce_topk_tc: needs k <= min(n, 256)");
  // ---- pass A: exact top-k over a sample (a prefix of the rows) -> tau ------------------------------
  const uint64_t sample = level == 0 ? std::min<uint64_t>(n, std::max<uint64_t>(16384, n / 32))
                                     : std::min<uint64_t>(n, std::max<uint64_t>(std::max<uint64_t>(2048, 4ull * k), n / 16));
  // grow-only workspace kept per (device, level): cudaMalloc / cudaFree around every call cost ~4 ms of the 16
  // (serialised by the lock taken in bruteforce_topk_tc)
  static TcWorkspace ws_all[16][2];
  int dev = 0;
  AB_CUDA(cudaGetDevice(&dev));
  TcWorkspace &W = ws_all[dev & 15][level];
  DevBuf<float> &sd = W.sd, &xnorm = W.xnorm, &qnorm = W.qnorm, &thr = W.thr, &rd = W.rd;
  DevBuf<uint32_t> &si = W.si, &maxbits = W.maxbits, &cand = W.cand, &counts = W.counts, &redo = W.redo, &ri = W.ri;
  sd.reserve(nq * k);
  si.reserve(nq * k);
  if (level == 0 && sample >= 16384 && sample < n) {
    tc_run(metric, d_rows, sample, dim, stride, d_queries, nq, qstride, k, si.p, sd.p, sm_count, s, 1);
  } else {
    bruteforce_topk(metric, d_rows, sample, stride, d_queries, nq, qstride, k, si.p, sd.p, sm_count, s);
  }

  xnorm.reserve(n);
  qnorm.reserve(nq);
  thr.reserve(nq);
  maxbits.alloc_zero(2, s);
  row_norms_kernel<<<sm_count * 8, 256, 0, s>>>(d_rows, n, stride, xnorm.p, maxbits.p);
  row_norms_kernel<<<std::max<int>(1, std::min<uint64_t>((nq * 32 + 255) / 256, uint64_t(sm_count) * 8)), 256, 0, s>>>(
      d_queries, nq, qstride, qnorm.p, maxbits.p + 1);
  AB_CUDA(cudaGetLastError());
  thresholds_kernel<<<static_cast<unsigned>((nq + 255) / 256), 256, 0, s>>>(sd.p, si.p, k, qnorm.p, maxbits.p, l2 ? 1 : 0, nq, thr.p);
  AB_CUDA(cudaGetLastError());
  if (!l2) AB_CUDA(cudaMemsetAsync(xnorm.p, 0, n * sizeof(float), s));  // IP: no per-row term
  count_launch(3);

  // ---- pass B: tcgen05 filter ---------------------------------------------------------------------
  const uint32_t qtiles = static_cast<uint32_t>((nq + TC_M - 1) / TC_M);
  const uint64_t total_tiles = (n + TC_N - 1) / TC_N;
  // one CTA per SM (197 KB of shared memory): pick the split count whose CTA total fills whole waves best
  uint32_t splits = 1;
  {
    double best = -1.0;
    const uint32_t smax = static_cast<uint32_t>(std::min<uint64_t>(total_tiles, std::max<uint32_t>(1, uint32_t(sm_count) * 8 / qtiles + 1)));
    for (uint32_t sp = 1; sp <= smax; sp++) {
      const uint64_t ctas = uint64_t(qtiles) * sp;
      const uint64_t waves = (ctas + sm_count - 1) / sm_count;
      const double eff = double(ctas) / double(waves * sm_count) - 0.002 * sp;  // mild preference for fewer splits
      if (ctas >= uint64_t(sm_count) && eff > best) {
        best = eff;
        splits = sp;
      }
    }
    if (best < 0) splits = smax;
  }
  const uint32_t tiles_per_split = static_cast<uint32_t>((total_tiles + splits - 1) / splits);
  uint64_t expect = uint64_t(k) * ((n + sample - 1) / sample);  // survivors expected per query over all splits
  uint32_t cap = 256;
  while (cap < 4 * expect / splits + 64 && cap < 16384) cap <<= 1;
  cand.reserve(size_t(splits) * nq * cap);
  counts.alloc_zero(size_t(splits) * nq, s);
  const CUtensorMap map_q = make_map(d_queries, nq, dim, qstride, TC_M);
  const CUtensorMap map_x = make_map(d_rows, n, dim, stride, TC_N);
  TcArgs a{};
  a.n = n;
  a.nq = nq;
  a.kblocks = (dim + TC_KB - 1) / TC_KB;
  a.tiles_per_split = tiles_per_split;
  a.alpha = l2 ? -2.f : -1.f;
  a.xnorm = xnorm.p;
  a.thr = thr.p;
  a.cand = cand.p;
  a.counts = counts.p;
  a.cap = cap;
  if (a.kblocks <= 4) {
    AB_CUDA(cudaFuncSetAttribute(bf_tc_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(TC_SMEM)));
    bf_tc_kernel<true><<<dim3(qtiles, splits), TC_THREADS, TC_SMEM, s>>>(map_q, map_x, a);
  } else {
    AB_CUDA(cudaFuncSetAttribute(bf_tc_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(TC_SMEM)));
    bf_tc_kernel<false><<<dim3(qtiles, splits), TC_THREADS, TC_SMEM, s>>>(map_q, map_x, a);
  }
  AB_CUDA(cudaGetLastError());
  count_launch();

  // ---- pass C: exact re-evaluation ----------------------------------------------------------------
  redo.reserve(nq + 1);
  AB_CUDA(cudaMemsetAsync(redo.p + nq, 0, 4, s));
  const uint32_t need = static_cast<uint32_t>(std::min<uint64_t>(k, n));
  const int nchunks = static_cast<int>(stride / 4);
  const int warps = 4;
  const size_t smem = round_up16(size_t(nchunks) * 16 + size_t(k) * 8 + 64 * 4) * warps;
  const int blocks = static_cast<int>(std::min<uint64_t>((nq + warps - 1) / warps, uint64_t(sm_count) * 16));
  if (l2) {
    AB_CUDA(cudaFuncSetAttribute(refine_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem)));
    refine_kernel<false><<<blocks, warps * 32, smem, s>>>(d_rows, stride, nchunks, d_queries, qstride, nq, cand.p, counts.p, splits, cap, k,
                                                       need, d_out_ids, d_out_dist, redo.p, redo.p + nq, level);
  } else {
    AB_CUDA(cudaFuncSetAttribute(refine_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem)));
    refine_kernel<true><<<blocks, warps * 32, smem, s>>>(d_rows, stride, nchunks, d_queries, qstride, nq, cand.p, counts.p, splits, cap, k,
                                                      need, d_out_ids, d_out_dist, redo.p, redo.p + nq, level);
  }
  AB_CUDA(cudaGetLastError());
  count_launch();

  // ---- queries whose survivor list overflowed: exact kernel ---------------------------------------------
  // The sample level needs no repair (see refine_kernel) and therefore no host round trip. The top level reads back
  // ONE counter -- the synchronisation the caller needs anyway -- and only a non-zero count costs a second one.
  if (level != 0) return;
  uint32_t n_redo = 0;
  AB_CUDA(cudaMemcpyAsync(&n_redo, redo.p + nq, 4, cudaMemcpyDeviceToHost, s));
  AB_CUDA(cudaStreamSynchronize(s));
  if (n_redo == 0) return;
  std::vector<uint32_t> flags(nq);
  AB_CUDA(cudaMemcpyAsync(flags.data(), redo.p, nq * 4, cudaMemcpyDeviceToHost, s));
  AB_CUDA(cudaStreamSynchronize(s));
  std::vector<uint32_t> which;
  for (uint64_t q = 0; q < nq; q++)
    if (flags[q]) which.push_back(static_cast<uint32_t>(q));
  const uint32_t m = static_cast<uint32_t>(which.size());
  DevBuf<uint32_t> dw;
  DevBuf<float> gq;
  dw.reserve(m);
  gq.reserve(size_t(m) * qstride);
  ri.reserve(size_t(m) * k);
  rd.reserve(size_t(m) * k);
  AB_CUDA(cudaMemcpyAsync(dw.p, which.data(), m * 4, cudaMemcpyHostToDevice, s));
  gather_rows_f32<<<std::min<uint64_t>((uint64_t(m) * qstride + 255) / 256, 4096), 256, 0, s>>>(d_queries, qstride, dw.p, m, gq.p);
  bruteforce_topk(metric, d_rows, n, stride, gq.p, m, qstride, k, ri.p, rd.p, sm_count, s);
  scatter_results<<<static_cast<unsigned>((uint64_t(m) * k + 255) / 256), 256, 0, s>>>(ri.p, rd.p, dw.p, m, k, d_out_ids, d_out_dist);
  AB_CUDA(cudaGetLastError());
  count_launch(2);
  AB_CUDA(cudaStreamSynchronize(s));
}

void bruteforce_topk_tc(int metric, const float *d_rows, uint64_t n, uint32_t dim, uint32_t stride, const float *d_queries,
                        uint64_t nq, uint32_t qstride, uint32_t k, uint32_t *d_out_ids, float *d_out_dist, int sm_count,
                        cudaStream_t s) {
  tc_run(metric, d_rows, n, dim, stride, d_queries, nq, qstride, k, d_out_ids, d_out_dist, sm_count, s, 0);
}

}  // namespace ab
