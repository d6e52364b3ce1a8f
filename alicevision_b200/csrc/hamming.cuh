// K3 — Hamming nearest-neighbour search for 64-byte binary descriptors (AKAZE-MLDB), no tensor cores.
// Replaces ArrayMatcher_bruteForce<unsigned char, Hamming<unsigned char>>::SearchNeighbours
// (matching/ArrayMatcher_bruteForce.hpp:98-142 with feature/Hamming.hpp:113-129: sum of popcount(a^b) over
// eight 64-bit words) + top-2 + the ratio test of matching/filters.hpp:60 on unsigned distances
// ((float)d1 < ratio * (float)d2, ratio NOT squared: matching/RegionsMatcher.cpp:163).
//
// One thread owns one query (16 x u32 in registers); the database image is streamed through shared memory in
// 256-row tiles (16 KB) and read with warp-broadcast LDS.128, so HBM/L2 traffic is one pass over the database
// per 256 queries.  16 XOR + 16 POPC + adds per (query,row): the kernel is bound by the integer/popc pipe.
#pragma once
#include "common.cuh"

namespace b200m {

constexpr int HM_TQ = 256;     // queries per block (one per thread)
constexpr int HM_TD = 256;     // database rows per smem tile

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
      uint32_t d = __popc(a0.x ^ b0.x) + __popc(a0.y ^ b0.y) + __popc(a0.z ^ b0.z) + __popc(a0.w ^ b0.w);
      d += __popc(a1.x ^ b1.x) + __popc(a1.y ^ b1.y) + __popc(a1.z ^ b1.z) + __popc(a1.w ^ b1.w);
      d += __popc(a2.x ^ b2.x) + __popc(a2.y ^ b2.y) + __popc(a2.z ^ b2.z) + __popc(a2.w ^ b2.w);
      d += __popc(a3.x ^ b3.x) + __popc(a3.y ^ b3.y) + __popc(a3.z ^ b3.z) + __popc(a3.w ^ b3.w);
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
