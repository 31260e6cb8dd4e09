// Device distance evaluation for the raw / SQ8 / SQ4 spaces.
//
// Replaces the reference's host SIMD kernels (paths relative to /root/reference):
//   simd::l2_sqr / ip_sqr            include/simd/distance_l2.ipp:41-181, distance_ip.ipp:45-170
//   simd::l2_sqr_sq8 / ip_sqr_sq8    distance_l2.ipp:187-408,  distance_ip.ipp:177-375
//   simd::l2_sqr_sq4 / ip_sqr_sq4    distance_l2.ipp:209-237,:413-676, distance_ip.ipp:380-686
//
// Layout: a row (vector or code string) is a sequence of 16-byte chunks; rows start on 64-byte
// boundaries (the reference's SequentialStorage stride) and are zero padded. A sub-warp of TPV
// lanes owns one candidate row: lane l reads chunks l, l+TPV, ... with 128-bit loads, accumulates
// its partial sum in fp32 with FMAs, and the partials are combined by an xor-butterfly.
// The query side lives in shared memory (raw floats or codes) next to per-dimension dequantisation
// tables (scale = (max-min)/levels, min), which are zero for padded dimensions so that padding
// never contributes.
//
// SQ4 nibble pairing follows the reference *distance kernels*, not its encoder: within byte j the
// LOW nibble is scored with table entry 2j and the HIGH nibble with entry 2j+1, although the
// encoder stored dimension 2j in the high nibble. Results parity
// requires reproducing that.
#pragma once
#include "common.cuh"

namespace ab {

struct QuerySmem {
  const float *q_raw;     // raw query, padded to a whole number of chunks (may be null for SQ traversal)
  const uint8_t *q_code;  // encoded query (SQ only)
  const float *scale;     // [table_len] per-dimension (max-min)*(1/levels), 0 in the padding
  const float *vmin;      // [table_len] per-dimension min, 0 in the padding (IP only)
};

// exact uint8 -> float without I2F: place the byte in the mantissa of 2^23
__device__ __forceinline__ float byte_as_magic(uint32_t w, int k) {
  return __uint_as_float(__byte_perm(w, 0x4B000000u, 0x7540u + k));
}
constexpr float kMagic = 8388608.0f;

template <bool IP>
__device__ __forceinline__ float raw_chunk(uint4 c, int chunk, const float *q_raw, float acc) {
  const float4 q = reinterpret_cast<const float4 *>(q_raw)[chunk];
  const float x0 = __uint_as_float(c.x), x1 = __uint_as_float(c.y), x2 = __uint_as_float(c.z),
              x3 = __uint_as_float(c.w);
  if (IP) {
    acc = fmaf(q.x, x0, acc);
    acc = fmaf(q.y, x1, acc);
    acc = fmaf(q.z, x2, acc);
    acc = fmaf(q.w, x3, acc);
  } else {
    float d0 = q.x - x0, d1 = q.y - x1, d2 = q.z - x2, d3 = q.w - x3;
    acc = fmaf(d0, d0, acc);
    acc = fmaf(d1, d1, acc);
    acc = fmaf(d2, d2, acc);
    acc = fmaf(d3, d3, acc);
  }
  return acc;
}

// one 32-bit word = 4 codes, table entries t[0..3]
template <bool IP>
__device__ __forceinline__ float sq_word(uint32_t cw, uint32_t qw, const float4 s, const float4 m, float acc) {
  const float sv[4] = {s.x, s.y, s.z, s.w};
  const float mv[4] = {m.x, m.y, m.z, m.w};
#pragma unroll
  for (int k = 0; k < 4; k++) {
    const float fc = byte_as_magic(cw, k), fq = byte_as_magic(qw, k);
    if (IP) {
      // -(sum (min + cq*s) * (min + cc*s))   distance_ip.ipp:177-191
      const float xv = fmaf(fq - kMagic, sv[k], mv[k]);
      const float yv = fmaf(fc - kMagic, sv[k], mv[k]);
      acc = fmaf(xv, yv, acc);
    } else {
      // sum ((cq - cc) * s)^2                 distance_l2.ipp:187-201
      const float sd = (fq - fc) * sv[k];
      acc = fmaf(sd, sd, acc);
    }
  }
  return acc;
}

template <bool IP>
__device__ __forceinline__ float sq8_chunk(uint4 c, int chunk, const QuerySmem &qs, float acc) {
  const uint4 q = reinterpret_cast<const uint4 *>(qs.q_code)[chunk];
  const float4 *s = reinterpret_cast<const float4 *>(qs.scale) + chunk * 4;
  const float4 *m = reinterpret_cast<const float4 *>(qs.vmin) + chunk * 4;
  const float4 z = make_float4(0.f, 0.f, 0.f, 0.f);
  acc = sq_word<IP>(c.x, q.x, s[0], IP ? m[0] : z, acc);
  acc = sq_word<IP>(c.y, q.y, s[1], IP ? m[1] : z, acc);
  acc = sq_word<IP>(c.z, q.z, s[2], IP ? m[2] : z, acc);
  acc = sq_word<IP>(c.w, q.w, s[3], IP ? m[3] : z, acc);
  return acc;
}

// one 32-bit word = 4 bytes = 8 codes; table entries 8w .. 8w+7 in the order lo0,hi0,lo1,hi1,...
template <bool IP>
__device__ __forceinline__ float sq4_word(uint32_t cw, uint32_t qw, const float4 *s, const float4 *m, float acc) {
  const uint32_t clo = cw & 0x0F0F0F0Fu, chi = (cw >> 4) & 0x0F0F0F0Fu;
  const uint32_t qlo = qw & 0x0F0F0F0Fu, qhi = (qw >> 4) & 0x0F0F0F0Fu;
  const float4 s0 = s[0], s1 = s[1];
  const float sv[8] = {s0.x, s0.y, s0.z, s0.w, s1.x, s1.y, s1.z, s1.w};
  float mv[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
  if (IP) {
    const float4 m0 = m[0], m1 = m[1];
    mv[0] = m0.x; mv[1] = m0.y; mv[2] = m0.z; mv[3] = m0.w;
    mv[4] = m1.x; mv[5] = m1.y; mv[6] = m1.z; mv[7] = m1.w;
  }
#pragma unroll
  for (int k = 0; k < 4; k++) {
#pragma unroll
    for (int h = 0; h < 2; h++) {
      const float fc = byte_as_magic(h ? chi : clo, k), fq = byte_as_magic(h ? qhi : qlo, k);
      const int t = 2 * k + h;
      if (IP) {
        const float xv = fmaf(fq - kMagic, sv[t], mv[t]);
        const float yv = fmaf(fc - kMagic, sv[t], mv[t]);
        acc = fmaf(xv, yv, acc);
      } else {
        const float sd = (fq - fc) * sv[t];
        acc = fmaf(sd, sd, acc);
      }
    }
  }
  return acc;
}

template <bool IP>
__device__ __forceinline__ float sq4_chunk(uint4 c, int chunk, const QuerySmem &qs, float acc) {
  const uint4 q = reinterpret_cast<const uint4 *>(qs.q_code)[chunk];
  const float4 *s = reinterpret_cast<const float4 *>(qs.scale) + chunk * 8;
  const float4 *m = reinterpret_cast<const float4 *>(qs.vmin) + chunk * 8;
  acc = sq4_word<IP>(c.x, q.x, s + 0, m + 0, acc);
  acc = sq4_word<IP>(c.y, q.y, s + 2, m + 2, acc);
  acc = sq4_word<IP>(c.z, q.z, s + 4, m + 4, acc);
  acc = sq4_word<IP>(c.w, q.w, s + 6, m + 6, acc);
  return acc;
}

template <int QUANT, bool IP>
__device__ __forceinline__ float chunk_term(uint4 c, int chunk, const QuerySmem &qs, float acc) {
  if (QUANT == AB_QUANT_NONE) return raw_chunk<IP>(c, chunk, qs.q_raw, acc);
  if (QUANT == AB_QUANT_SQ8) return sq8_chunk<IP>(c, chunk, qs, acc);
  return sq4_chunk<IP>(c, chunk, qs, acc);
}

// Partial distance of one row for this lane of a TPV-wide group, then group reduction.
// Every lane of the group returns the full value. `row` may be any address when !active.
template <int QUANT, bool IP, int TPV, int MAXC>
__device__ __forceinline__ float row_distance(const uint8_t *row, int nchunks, int lane_in_group,
                                              const QuerySmem &qs, bool active) {
  uint4 c[MAXC];
#pragma unroll
  for (int i = 0; i < MAXC; i++) {
    const int ch = lane_in_group + i * TPV;
    c[i] = (active && ch < nchunks) ? ldg128(row + size_t(ch) * 16) : make_uint4(0, 0, 0, 0);
  }
  float acc = 0.f;
#pragma unroll
  for (int i = 0; i < MAXC; i++) {
    const int ch = lane_in_group + i * TPV;
    if (active && ch < nchunks) acc = chunk_term<QUANT, IP>(c[i], ch, qs, acc);
  }
  // rows longer than MAXC*TPV chunks: stream the remainder
  for (int ch = lane_in_group + MAXC * TPV; ch < nchunks; ch += TPV) {
    if (active) acc = chunk_term<QUANT, IP>(ldg128(row + size_t(ch) * 16), ch, qs, acc);
  }
#pragma unroll
  for (int o = TPV / 2; o > 0; o >>= 1) acc += __shfl_xor_sync(0xFFFFFFFFu, acc, o);
  return IP ? -acc : acc;
}

// ---- raw-space fast path --------------------------------------------------------------------------
// The query slice of a lane lives in registers (CPL chunks per lane, lane l of a TPV-wide group owns
// chunks l, l+TPV, ...), row chunks arrive as two packed-f32x2 halves per 128-bit load, and the
// arithmetic uses Blackwell's packed fp32 pipe (sub.f32x2 / fma.rn.f32x2 -> FADD2 / FFMA2): 4 FP
// instructions per 16-byte chunk instead of 8. Two candidates per group are in flight at once, so a
// lane has up to 2*CPL independent 128-bit loads outstanding before the first FMA. The loads and the
// packed arithmetic are volatile asm on purpose: ptxas otherwise sinks the loads next to their uses
// to save registers, which serialises the round trips (seen in the r01 SASS: 2 loads in flight, not 8).
struct U64x2 {
  unsigned long long a, b;
};
__device__ __forceinline__ U64x2 ldg128_pairs(const void *p) {
  U64x2 r;
  // plain non-coherent load: measured against L1::no_allocate, ld.global.cg and L1::evict_first on the bench
  // workload (r01: no_allocate -4 %, evict_first -10 %, cg +-0)
  asm volatile("ld.global.nc.v2.u64 {%0,%1}, [%2];" : "=l"(r.a), "=l"(r.b) : "l"(p));
  return r;
}
__device__ __forceinline__ unsigned long long f2sub(unsigned long long x, unsigned long long y) {
  unsigned long long d;
  asm volatile("sub.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(x), "l"(y));
  return d;
}
__device__ __forceinline__ unsigned long long f2fma(unsigned long long x, unsigned long long y, unsigned long long z) {
  unsigned long long d;
  asm volatile("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(x), "l"(y), "l"(z));
  return d;
}
__device__ __forceinline__ float f2sum(unsigned long long x) {
  return __uint_as_float(static_cast<uint32_t>(x)) + __uint_as_float(static_cast<uint32_t>(x >> 32));
}

template <bool IP>
__device__ __forceinline__ unsigned long long raw_pairs_term(const U64x2 &q, const U64x2 &x, unsigned long long acc) {
  if (IP) {
    acc = f2fma(q.a, x.a, acc);
    acc = f2fma(q.b, x.b, acc);
  } else {
    const unsigned long long d0 = f2sub(q.a, x.a), d1 = f2sub(q.b, x.b);
    acc = f2fma(d0, d0, acc);
    acc = f2fma(d1, d1, acc);
  }
  return acc;
}

// load this lane's query slice from the staged (zero padded) query in shared memory
template <int TPV, int CPL, bool EXACT>
__device__ __forceinline__ void load_query_regs(const float *q_smem, int nchunks, int lg, U64x2 (&q)[CPL]) {
#pragma unroll
  for (int i = 0; i < CPL; i++) {
    const int ch = lg + i * TPV;
    if (EXACT || ch < nchunks) {
      const ulonglong2 t = reinterpret_cast<const ulonglong2 *>(q_smem)[ch];
      q[i].a = t.x;
      q[i].b = t.y;
    } else {
      q[i].a = 0ull;
      q[i].b = 0ull;
    }
  }
}

// rows r0 / r1 already point at this lane's first chunk (row + lg * 16)
template <bool IP, int TPV, int CPL, bool EXACT>
__device__ __forceinline__ void raw_distance2(const uint8_t *r0, const uint8_t *r1, const U64x2 (&q)[CPL], int nchunks, int lg,
                                              float &d) {
  U64x2 x0[CPL], x1[CPL];
#pragma unroll
  for (int i = 0; i < CPL; i++) {
    if (EXACT || lg + i * TPV < nchunks) {
      x0[i] = ldg128_pairs(r0 + i * TPV * 16);
      x1[i] = ldg128_pairs(r1 + i * TPV * 16);
    } else {
      x0[i].a = x0[i].b = x1[i].a = x1[i].b = 0ull;
    }
  }
  __syncwarp();  // scheduling fence: keeps ptxas from sinking the last loads below the first arithmetic
  unsigned long long a0 = 0ull, a1 = 0ull;
#pragma unroll
  for (int i = 0; i < CPL; i++) {
    a0 = raw_pairs_term<IP>(q[i], x0[i], a0);
    a1 = raw_pairs_term<IP>(q[i], x1[i], a1);
  }
  // two sums across TPV lanes in log2(TPV) + 1 shuffles: the first exchange hands each lane the half it does
  // not keep (even lanes keep candidate 0, odd lanes candidate 1), then one butterfly over the remaining bits
  const float s0 = f2sum(a0), s1 = f2sum(a1);
  const bool odd = lg & 1;
  float s = (odd ? s1 : s0) + __shfl_xor_sync(0xFFFFFFFFu, odd ? s0 : s1, 1);
#pragma unroll
  for (int o = TPV / 2; o > 1; o >>= 1) s += __shfl_xor_sync(0xFFFFFFFFu, s, o);
  // even lanes of the group now hold candidate 0's distance, odd lanes candidate 1's
  d = IP ? -s : s;
}

template <bool IP, int TPV, int CPL, bool EXACT>
__device__ __forceinline__ float raw_distance1(const uint8_t *r0, const U64x2 (&q)[CPL], int nchunks, int lg) {
  U64x2 x0[CPL];
#pragma unroll
  for (int i = 0; i < CPL; i++) {
    if (EXACT || lg + i * TPV < nchunks) {
      x0[i] = ldg128_pairs(r0 + i * TPV * 16);
    } else {
      x0[i].a = x0[i].b = 0ull;
    }
  }
  unsigned long long a0 = 0ull;
#pragma unroll
  for (int i = 0; i < CPL; i++) a0 = raw_pairs_term<IP>(q[i], x0[i], a0);
  float s0 = f2sum(a0);
#pragma unroll
  for (int o = TPV / 2; o > 0; o >>= 1) s0 += __shfl_xor_sync(0xFFFFFFFFu, s0, o);
  return IP ? -s0 : s0;
}

// Sum three (two, one) per-lane partials over the TPV lanes of a group at once: a transposed butterfly leaves
// candidate c's total in the lanes with (lg % P) == c, P = 1 / 2 / 4 for NC = 1 / 2 / 3, in log2(TPV) + NC - 1 shuffles.
template <bool IP, int TPV, int NC>
__device__ __forceinline__ float reduce_multi(float s, float s1, float s2, int lg) {
  int first = 1;  // first butterfly offset still to do
  if (NC == 2) {
    const bool b0 = lg & 1;
    s = (b0 ? s1 : s) + __shfl_xor_sync(0xFFFFFFFFu, b0 ? s : s1, 1);
    first = 2;
  } else if (NC == 3) {
    const bool b0 = lg & 1, b1 = lg & 2;
    // xor 1: even lanes keep candidates {0, 2}, odd lanes {1, pad}
    const float t0 = (b0 ? s1 : s) + __shfl_xor_sync(0xFFFFFFFFu, b0 ? s : s1, 1);
    const float t1 = (b0 ? 0.f : s2) + __shfl_xor_sync(0xFFFFFFFFu, b0 ? s2 : 0.f, 1);
    // xor 2: lanes with bit 1 clear keep the first of their pair, the others the second
    s = (b1 ? t1 : t0) + __shfl_xor_sync(0xFFFFFFFFu, b1 ? t0 : t1, 2);
    first = 4;
  }
#pragma unroll
  for (int o = TPV / 2; o >= 1; o >>= 1) {
    if (o >= first) s += __shfl_xor_sync(0xFFFFFFFFu, s, o);
  }
  return IP ? -s : s;
}

// Up to three candidates per group in flight with the query chunks read from shared memory *after* the
// row loads have been issued: no registers are pinned by the query while the loads are outstanding, so
// a lane can keep 3 * CPL 128-bit loads in flight inside the same register budget (12 at d = 128).
// Returns candidate c's distance in the lanes with (lg % P) == c, P = 1 / 2 / 4 for NC = 1 / 2 / 3.
template <bool IP, int TPV, int CPL, bool EXACT, int NC>
__device__ __forceinline__ float raw_distance_multi(const uint8_t *r0, const uint8_t *r1, const uint8_t *r2,
                                                    const float *q_smem, int nchunks, int lg) {
  static_assert(NC >= 1 && NC <= 3 && TPV >= 4, "");
  U64x2 x0[CPL], x1[NC > 1 ? CPL : 1], x2[NC > 2 ? CPL : 1];
#pragma unroll
  for (int i = 0; i < CPL; i++) {
    if (EXACT || lg + i * TPV < nchunks) {
      x0[i] = ldg128_pairs(r0 + i * TPV * 16);
      if (NC > 1) x1[i] = ldg128_pairs(r1 + i * TPV * 16);
      if (NC > 2) x2[i] = ldg128_pairs(r2 + i * TPV * 16);
    } else {
      x0[i].a = x0[i].b = 0ull;
      if (NC > 1) x1[i].a = x1[i].b = 0ull;
      if (NC > 2) x2[i].a = x2[i].b = 0ull;
    }
  }
  __syncwarp();  // scheduling fence: every load is issued before the first shared-memory read / FMA
  unsigned long long a0 = 0ull, a1 = 0ull, a2 = 0ull;
#pragma unroll
  for (int i = 0; i < CPL; i++) {
    U64x2 q;
    q.a = q.b = 0ull;
    if (EXACT || lg + i * TPV < nchunks) {
      const ulonglong2 t = reinterpret_cast<const ulonglong2 *>(q_smem)[lg + i * TPV];
      q.a = t.x;
      q.b = t.y;
    }
    a0 = raw_pairs_term<IP>(q, x0[i], a0);
    if (NC > 1) a1 = raw_pairs_term<IP>(q, x1[i], a1);
    if (NC > 2) a2 = raw_pairs_term<IP>(q, x2[i], a2);
  }
  return reduce_multi<IP, TPV, NC>(f2sum(a0), f2sum(a1), f2sum(a2), lg);
}

// Six (five, four) candidates per group in flight: the deep variant for launches that shared memory limits to a
// dozen resident warps (large ef). The registers that twice the warps would have used hold twice the rows instead,
// so a hop of up to 24 new neighbours is evaluated in ONE pass over memory instead of two. Candidates 0..2 come back
// like raw_distance_multi<3> (lanes with lg % 4 == c), candidates 3..5 in `d_hi` the same way.
template <bool IP, int TPV, int CPL, bool EXACT, int NC>
__device__ __forceinline__ float raw_distance_deep(const uint8_t *const (&r)[6], const float *q_smem, int nchunks, int lg, float &d_hi) {
  static_assert(NC >= 4 && NC <= 6 && TPV >= 4, "");
  U64x2 x[NC][CPL];
#pragma unroll
  for (int i = 0; i < CPL; i++) {
#pragma unroll
    for (int c = 0; c < NC; c++) {
      if (EXACT || lg + i * TPV < nchunks) {
        x[c][i] = ldg128_pairs(r[c] + i * TPV * 16);
      } else {
        x[c][i].a = x[c][i].b = 0ull;
      }
    }
  }
  __syncwarp();  // scheduling fence: every load is issued before the first shared-memory read / FMA
  unsigned long long a[6] = {0ull, 0ull, 0ull, 0ull, 0ull, 0ull};
#pragma unroll
  for (int i = 0; i < CPL; i++) {
    U64x2 q;
    q.a = q.b = 0ull;
    if (EXACT || lg + i * TPV < nchunks) {
      const ulonglong2 t = reinterpret_cast<const ulonglong2 *>(q_smem)[lg + i * TPV];
      q.a = t.x;
      q.b = t.y;
    }
#pragma unroll
    for (int c = 0; c < NC; c++) a[c] = raw_pairs_term<IP>(q, x[c][i], a[c]);
  }
  d_hi = reduce_multi<IP, TPV, NC - 3>(f2sum(a[3]), f2sum(a[4]), f2sum(a[5]), lg);
  return reduce_multi<IP, TPV, 3>(f2sum(a[0]), f2sum(a[1]), f2sum(a[2]), lg);
}

// ---- SQ8 / SQ4, three candidates per group in flight ----------------------------------------------
// Same shape as raw_distance_multi. A lane owns chunk index lg + i * TPV of all three candidates, so the
// query codes and the per-dimension tables of a chunk are expanded once and used three times, and the
// arithmetic runs on pairs of dimensions in the packed fp32 pipe: per code ~2.9 instructions instead of 5
// (SQ8) / ~3.2 instead of 5.5 (SQ4). Each term is computed exactly as in sq_word / sq4_word (and in the
// reference kernels, distance_l2.ipp:187-201, distance_ip.ipp:177-191); only the summation order differs.
__device__ __forceinline__ unsigned long long p2pack(float lo, float hi) {
  unsigned long long r;
  asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(lo), "f"(hi));
  return r;
}
__device__ __forceinline__ unsigned long long p2sub(unsigned long long x, unsigned long long y) {
  unsigned long long d;
  asm("sub.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(x), "l"(y));
  return d;
}
__device__ __forceinline__ unsigned long long p2mul(unsigned long long x, unsigned long long y) {
  unsigned long long d;
  asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(x), "l"(y));
  return d;
}
__device__ __forceinline__ unsigned long long p2fma(unsigned long long x, unsigned long long y, unsigned long long z) {
  unsigned long long d;
  asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(x), "l"(y), "l"(z));
  return d;
}

// NP pairs of magic floats (2^23 + code) out of one 32-bit word of codes, in table order:
// SQ8: (b0,b1) (b2,b3); SQ4: (lo_k, hi_k) for the four bytes k (table entries 2k, 2k+1, see sq4_word)
template <int QUANT>
__device__ __forceinline__ void word_pairs(uint32_t w, unsigned long long (&out)[QUANT == AB_QUANT_SQ8 ? 2 : 4]) {
  if (QUANT == AB_QUANT_SQ8) {
    out[0] = p2pack(byte_as_magic(w, 0), byte_as_magic(w, 1));
    out[1] = p2pack(byte_as_magic(w, 2), byte_as_magic(w, 3));
  } else {
    // the high nibbles stay where they are (16 x code): the staged scale table holds s / 16 for the odd entries, so
    // (16 c) * (s / 16) is the same real product as c * s and every term rounds exactly as before, one shift less
    const uint32_t lo = w & 0x0F0F0F0Fu, hi = w & 0xF0F0F0F0u;
#pragma unroll
    for (int k = 0; k < 4; k++) out[k] = p2pack(byte_as_magic(lo, k), byte_as_magic(hi, k));
  }
}

template <int QUANT, bool IP, int TPV, int CPL, int NC>
__device__ __forceinline__ float sq_distance_multi(const uint8_t *r0, const uint8_t *r1, const uint8_t *r2, const QuerySmem &qs,
                                                   int nchunks, int lg) {
  static_assert(NC >= 1 && NC <= 3 && TPV >= 4 && QUANT != AB_QUANT_NONE, "");
  constexpr int NP = QUANT == AB_QUANT_SQ8 ? 2 : 4;  // pairs of dimensions per 32-bit word of codes
  constexpr int kUnrollWords = QUANT == AB_QUANT_SQ8 ? 4 : 1;
  uint4 x0[CPL], x1[NC > 1 ? CPL : 1], x2[NC > 2 ? CPL : 1];
#pragma unroll
  for (int i = 0; i < CPL; i++) {
    if (lg + i * TPV < nchunks) {
      x0[i] = ldg128(r0 + i * TPV * 16);
      if (NC > 1) x1[i] = ldg128(r1 + i * TPV * 16);
      if (NC > 2) x2[i] = ldg128(r2 + i * TPV * 16);
    } else {
      x0[i] = make_uint4(0, 0, 0, 0);
      if (NC > 1) x1[i] = make_uint4(0, 0, 0, 0);
      if (NC > 2) x2[i] = make_uint4(0, 0, 0, 0);
    }
  }
  __syncwarp();  // scheduling fence: every load is issued before the first shared-memory read / FMA
  const unsigned long long magic2 = p2pack(kMagic, kMagic);
  unsigned long long a0 = 0ull, a1 = 0ull, a2 = 0ull;
#pragma unroll
  for (int i = 0; i < CPL; i++) {
    const int ch = lg + i * TPV;
    if (ch < nchunks) {
      // SQ4: the four 32-bit words of the chunk go through one rolled loop (the vectors rotate by one word
      // per iteration): fully unrolled, that kernel is 14k instructions and spends half its time on
      // instruction fetch. SQ8 is half the size and is faster unrolled.
      uint4 q = reinterpret_cast<const uint4 *>(qs.q_code)[ch];
      uint4 c0 = x0[i], c1 = x1[NC > 1 ? i : 0], c2 = x2[NC > 2 ? i : 0];
      // table entries of this chunk: 16 (SQ8) or 32 (SQ4) dimensions = NP * 2 float4, stored transposed
      // ([float4 index within the chunk][chunk], see the staging loop of beam_search_kernel)
      const float4 *st = reinterpret_cast<const float4 *>(qs.scale) + ch;
      const float4 *mt = reinterpret_cast<const float4 *>(qs.vmin) + ch;
#pragma unroll(kUnrollWords)
      for (int w = 0; w < 4; w++) {
        unsigned long long S[NP], M[NP], Q[NP], C[NP];
#pragma unroll
        for (int h = 0; h < NP / 2; h++) {
          const float4 sv = st[h * nchunks];
          S[2 * h] = p2pack(sv.x, sv.y);
          S[2 * h + 1] = p2pack(sv.z, sv.w);
          if (IP) {
            const float4 mv = mt[h * nchunks];
            M[2 * h] = p2pack(mv.x, mv.y);
            M[2 * h + 1] = p2pack(mv.z, mv.w);
          }
        }
        st += (NP / 2) * nchunks;
        mt += (NP / 2) * nchunks;
        // SQ8 L2: |cq - cc| of four codes in ONE instruction (VABSDIFF4 is native on sm_100a), and the 2^23 bias of
        // the byte -> float trick folded into the product: fma(2^23 + a, s, -2^23 s) = fl(a * s) exactly (2^23 s is
        // exact), which is |fl((cq - cc) * s)| -- the reference's term up to the sign that the square removes. No
        // query-side expansion, no separate subtraction: 9 instead of ~11.3 instructions per word and candidate.
        constexpr bool ABSDIFF = QUANT == AB_QUANT_SQ8 && !IP;
        if (ABSDIFF) {
#pragma unroll
          for (int p = 0; p < NP; p++) M[p] = p2mul(S[p], p2pack(-kMagic, -kMagic));  // -2^23 * s
        } else {
          word_pairs<QUANT>(q.x, Q);
        }
        if (IP) {
#pragma unroll
          for (int p = 0; p < NP; p++) Q[p] = p2fma(p2sub(Q[p], magic2), S[p], M[p]);  // min + cq * s
        }
#pragma unroll
        for (int c = 0; c < NC; c++) {
          const uint32_t cw = c == 0 ? c0.x : (c == 1 ? c1.x : c2.x);
          word_pairs<QUANT>(ABSDIFF ? __vabsdiffu4(cw, q.x) : cw, C);
          unsigned long long acc = c == 0 ? a0 : (c == 1 ? a1 : a2);
#pragma unroll
          for (int p = 0; p < NP; p++) {
            if (IP) {
              acc = p2fma(Q[p], p2fma(p2sub(C[p], magic2), S[p], M[p]), acc);  // (min + cq*s) * (min + cc*s)
            } else if (ABSDIFF) {
              const unsigned long long sd = p2fma(C[p], S[p], M[p]);  // |cq - cc| * s
              acc = p2fma(sd, sd, acc);
            } else {
              const unsigned long long sd = p2mul(p2sub(Q[p], C[p]), S[p]);  // (cq - cc) * s
              acc = p2fma(sd, sd, acc);
            }
          }
          if (c == 0) a0 = acc; else if (c == 1) a1 = acc; else a2 = acc;
        }
        q = make_uint4(q.y, q.z, q.w, 0u);
        c0 = make_uint4(c0.y, c0.z, c0.w, 0u);
        if (NC > 1) c1 = make_uint4(c1.y, c1.z, c1.w, 0u);
        if (NC > 2) c2 = make_uint4(c2.y, c2.z, c2.w, 0u);
      }
    }
  }
  return reduce_multi<IP, TPV, NC>(f2sum(a0), f2sum(a1), f2sum(a2), lg);
}

// ---- quantizer (include/space/quant/sq8.hpp:120-132, sq4.hpp:120-132) --------------------------
// IEEE round-to-nearest sub / div / mul with no contraction, then truncation, so that codes equal
// the host's static_cast<uint8_t>(((float)v - min) / (max - min) * levels).
__device__ __forceinline__ uint32_t sq_quantize(float v, float mn, float mx, float levels) {
  if (mx == mn) return 0u;
  if (v >= mx) return static_cast<uint32_t>(levels);
  if (v <= mn) return 0u;
  const float scaled = __fdiv_rn(__fsub_rn(v, mn), __fsub_rn(mx, mn));
  return static_cast<uint32_t>(__fmul_rn(scaled, levels)) & 0xFFu;
}

}  // namespace ab
