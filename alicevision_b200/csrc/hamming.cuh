// K3 — Hamming nearest-neighbour search for 64-byte binary descriptors (AKAZE-MLDB), no tensor cores.
// Replaces ArrayMatcher_bruteForce<unsigned char, Hamming<unsigned char>>::SearchNeighbours
// (matching/ArrayMatcher_bruteForce.hpp:98-142 with feature/Hamming.hpp:113-129: sum of popcount(a^b) over
// eight 64-bit words) + top-2 + the ratio test of matching/filters.hpp:60 on unsigned distances
// ((float)d1 < ratio * (float)d2, ratio NOT squared: matching/RegionsMatcher.cpp:163).
//
// One thread owns one query (16 x u32 in registers); the database image is streamed through shared memory in
// 256-row tiles (16 KB) and read with warp-broadcast LDS.128, so HBM/L2 traffic is one pass over the database
// per 256 queries.  The first version (16 XOR + 16 POPC per (query,row)) ran the POPC (XU) pipe at 97.7 % (ncu,
// profiles/r01_ncu_hamming_top2_kernel.md) with the ALU pipe at 42 %; the 16 XOR words are now folded with a
// carry-save adder tree (11 full adders = 22 LOP3) down to five words of weight 1,1,2,4,8, so only 5 POPC remain.
#pragma once
#include "common.cuh"

namespace b200m {

constexpr int HM_TQ = 256;     // queries per block (one per thread)
constexpr int HM_TD = 256;     // database rows per smem tile

// full adder on 32 bit-columns at once: l = a^b^c (weight 1), h = majority(a,b,c) (weight 2)
__device__ __forceinline__ void csa(uint32_t& h, uint32_t& l, uint32_t a, uint32_t b, uint32_t c) {
  asm("lop3.b32 %0, %1, %2, %3, 0x96;" : "=r"(l) : "r"(a), "r"(b), "r"(c));
  asm("lop3.b32 %0, %1, %2, %3, 0xE8;" : "=r"(h) : "r"(a), "r"(b), "r"(c));
}
// popcount of 16 words = sum of popc over 512 bits, exact (feature/Hamming.hpp:120-129 sums eight popcountll)
__device__ __forceinline__ uint32_t popc512(const uint4& a0, const uint4& a1, const uint4& a2, const uint4& a3,
                                            const uint4& b0, const uint4& b1, const uint4& b2, const uint4& b3) {
  const uint32_t x0 = a0.x ^ b0.x, x1 = a0.y ^ b0.y, x2 = a0.z ^ b0.z, x3 = a0.w ^ b0.w;
  const uint32_t x4 = a1.x ^ b1.x, x5 = a1.y ^ b1.y, x6 = a1.z ^ b1.z, x7 = a1.w ^ b1.w;
  const uint32_t x8 = a2.x ^ b2.x, x9 = a2.y ^ b2.y, x10 = a2.z ^ b2.z, x11 = a2.w ^ b2.w;
  const uint32_t x12 = a3.x ^ b3.x, x13 = a3.y ^ b3.y, x14 = a3.z ^ b3.z, x15 = a3.w ^ b3.w;
  uint32_t c0, c1, c2, c3, c4, s0, s1, s2, s3, s4;           // level 1: 15 words -> 5 ones + 5 twos
  csa(c0, s0, x0, x1, x2); csa(c1, s1, x3, x4, x5); csa(c2, s2, x6, x7, x8); csa(c3, s3, x9, x10, x11); csa(c4, s4, x12, x13, x14);
  uint32_t d0, d1, t0, t1;                                   // ones: s0..s4, x15 -> t0, t1 (+ twos d0, d1)
  csa(d0, t0, s0, s1, s2); csa(d1, t1, s3, s4, x15);
  uint32_t e0, e1, u0, u1;                                   // twos: c0..c4, d0, d1 -> u0, u1, d1 (+ fours e0, e1)
  csa(e0, u0, c0, c1, c2); csa(e1, u1, c3, c4, d0);
  uint32_t f, v, g, w;
  csa(f, v, u0, u1, d1);                                     // twos -> v (+ four f)
  csa(g, w, f, e0, e1);                                      // fours -> w (+ eight g)
  return __popc(t0) + __popc(t1) + 2u * __popc(v) + 4u * __popc(w) + 8u * __popc(g);
}

template <bool OUT_DENSE>
__global__ void __launch_bounds__(HM_TQ)
hamming_top2_kernel(const ViewDev* __restrict__ views, const PairDev* __restrict__ pairs, Cand* __restrict__ cands,
                    int* __restrict__ cand_count, float ratio, int32_t* __restrict__ dense_idx, uint32_t* __restrict__ dense_dist) {
  const PairDev p = pairs[blockIdx.y];
  if (p.mode != PM_HAMMING) return;
  const int q = blockIdx.x * HM_TQ + threadIdx.x;
  if (blockIdx.x * HM_TQ >= (int)p.m_j) return;
  __shared__ uint4 tile[HM_TD * 4];
  const uint4* qraw = reinterpret_cast<const uint4*>(views[p.view_j].raw);
  const uint4* draw = reinterpret_cast<const uint4*>(views[p.view_i].raw);
  uint4 a0 = make_uint4(0, 0, 0, 0), a1 = a0, a2 = a0, a3 = a0;
  if (q < (int)p.m_j) { a0 = qraw[(size_t)q * 4]; a1 = qraw[(size_t)q * 4 + 1]; a2 = qraw[(size_t)q * 4 + 2]; a3 = qraw[(size_t)q * 4 + 3]; }
  uint32_t m1 = 0xFFFFFFFFu, m2 = 0xFFFFFFFFu; int i1 = 0x7fffffff, i2 = 0x7fffffff;
  for (int d0 = 0; d0 < (int)p.m_i; d0 += HM_TD) {
    __syncthreads();
    const int rows = min(HM_TD, (int)p.m_i - d0);
    for (int e = threadIdx.x; e < rows * 4; e += HM_TQ) tile[e] = draw[(size_t)d0 * 4 + e];
    __syncthreads();
#pragma unroll 4
    for (int r = 0; r < rows; ++r) {
      const uint4 b0 = tile[r * 4], b1 = tile[r * 4 + 1], b2 = tile[r * 4 + 2], b3 = tile[r * 4 + 3];
      const uint32_t d = popc512(a0, a1, a2, a3, b0, b1, b2, b3);
      // rows arrive in increasing index order, so strict '<' keeps (value, index) lexicographic order
      if (d < m2) {
        if (d < m1) { m2 = m1; i2 = i1; m1 = d; i1 = d0 + r; }
        else { m2 = d; i2 = d0 + r; }
      }
    }
  }
  if (q >= (int)p.m_j) return;
  if (OUT_DENSE) {
    dense_idx[2 * q] = i1; dense_idx[2 * q + 1] = i2;
    dense_dist[2 * q] = m1; dense_dist[2 * q + 1] = m2;
  } else if ((float)m1 < __fmul_rn(ratio, (float)m2)) {
    const int slot = atomicAdd(&cand_count[blockIdx.y], 1);
    cands[p.cand_base + slot] = Cand{(uint32_t)q, (uint32_t)i1, __uint_as_float(m1), __uint_as_float(m2)};
  }
}

}  // namespace b200m
