// Exactness pass + packing of the per-pair candidate lists into one contiguous record array.
//
// For tensor-core pairs (PM_TC) a candidate carries the exact best distance d1, the 16-row chunk of the
// database image that contains the best row, and an UPPER BOUND of the second-best distance (second
// smallest chunk minimum).  One warp re-scores the 16 rows of that chunk against the query with exact
// arithmetic (integer-valued fp16 data, fp32 accumulation of integers < 2^24), which yields
//   * the index of the nearest neighbour (first minimum inside the chunk),
//   * the true second-best distance min(bound, second minimum inside the chunk),
// and re-applies the reference's ratio test  d1 < ratio^2 * d2  (matching/filters.hpp:60,
// matching/RegionsMatcher.hpp:150) in float.  A mismatch between the re-scored and the tensor-core d1
// increments `err_count` (it would mean the tensor-core accumulation was not exact).
// For the other modes candidates are already exact and are only moved to their packed position.
#pragma once
#include "common.cuh"

namespace b200m {

// offsets[p] = exclusive prefix sum of cand_count[0..n); offsets[n] = total. Single block.
__global__ void scan_counts_kernel(const int* __restrict__ cand_count, int n, int* __restrict__ offsets) {
  __shared__ int carry;
  __shared__ int wsum[32];
  if (threadIdx.x == 0) carry = 0;
  __syncthreads();
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nw = blockDim.x >> 5;
  for (int base = 0; base < n; base += blockDim.x) {
    const int i = base + threadIdx.x;
    const int v = i < n ? cand_count[i] : 0;
    int x = v;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) { int y = __shfl_up_sync(0xffffffffu, x, o); if (lane >= o) x += y; }
    if (lane == 31) wsum[warp] = x;
    __syncthreads();
    if (warp == 0) {
      int s = lane < nw ? wsum[lane] : 0;
#pragma unroll
      for (int o = 1; o < 32; o <<= 1) { int y = __shfl_up_sync(0xffffffffu, s, o); if (lane >= o) s += y; }
      wsum[lane] = s;
    }
    __syncthreads();
    const int before = carry + (warp ? wsum[warp - 1] : 0) + x - v;
    if (i < n) offsets[i] = before;
    __syncthreads();
    if (threadIdx.x == blockDim.x - 1) carry = before + v;
    __syncthreads();
  }
  if (threadIdx.x == 0) offsets[n] = carry;
}

// Warp-level exact re-scoring of one tensor-core candidate (see the header comment). All 32 lanes call it with the same
// candidate; lane l handles database row 16*k.b + (l & 15), components 64*(l >> 4)..+63. Returns (on every lane)
// whether the match survives the ratio test; `rec` is the final record (i = database row, j = query row, d1, d2).
__device__ __forceinline__ bool rescore_candidate(const __half* __restrict__ db16, const __half* __restrict__ q16, uint32_t m_i, const Cand& k,
                                                  float ratio_sq, int lane, unsigned int* __restrict__ err_count, Rec& rec) {
  const int r = lane & 15, hv = lane >> 4;
  const uint32_t row = k.b * 16 + r;
  float acc = INFINITY;
  if (row < m_i) {
    const uint4* a = reinterpret_cast<const uint4*>(q16 + (size_t)k.q * 128 + hv * 64);
    const uint4* b = reinterpret_cast<const uint4*>(db16 + (size_t)row * 128 + hv * 64);
    acc = 0.f;
#pragma unroll
    for (int v = 0; v < 8; ++v) {
      const uint4 x = a[v], y = b[v];
      const __half2* xh = reinterpret_cast<const __half2*>(&x);
      const __half2* yh = reinterpret_cast<const __half2*>(&y);
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        const float2 fx = __half22float2(xh[e]), fy = __half22float2(yh[e]);
        const float d0 = fx.x - fy.x, d1 = fx.y - fy.y;
        acc = fmaf(d0, d0, acc);
        acc = fmaf(d1, d1, acc);
      }
    }
  }
  const float dist = acc + __shfl_xor_sync(0xffffffffu, acc, 16);      // INF for rows past the end
  float best = dist; int arg = r;                                        // argmin over the 16 rows (ties -> smallest row)
#pragma unroll
  for (int o = 8; o >= 1; o >>= 1) {
    const float ob = __shfl_xor_sync(0xffffffffu, best, o);
    const int oa = __shfl_xor_sync(0xffffffffu, arg, o);
    if (ob < best || (ob == best && oa < arg)) { best = ob; arg = oa; }
  }
  float second = (r == arg) ? INFINITY : dist;
#pragma unroll
  for (int o = 8; o >= 1; o >>= 1) second = fminf(second, __shfl_xor_sync(0xffffffffu, second, o));
  if (lane == 0 && best != k.d1) atomicAdd(err_count, 1u);             // tensor-core accumulation was not exact
  const float d2 = fminf(k.d2, second);
  rec = Rec{k.b * 16 + (uint32_t)arg, k.q, best, d2};
  return best < __fmul_rn(ratio_sq, d2);                                // matching/filters.hpp:60
}

constexpr int VERIFY_BLOCKS_PER_PAIR = 4;
constexpr int VERIFY_WARPS = 8;

__global__ void __launch_bounds__(VERIFY_WARPS * 32)
verify_pack_kernel(const ViewDev* __restrict__ views, const PairDev* __restrict__ pairs, const Cand* __restrict__ cands,
                   const int* __restrict__ cand_count, const int* __restrict__ offsets, Rec* __restrict__ out, float ratio_sq,
                   unsigned int* __restrict__ err_count) {
  const PairDev p = pairs[blockIdx.x];
  const int n = cand_count[blockIdx.x];
  const int lane = threadIdx.x & 31;
  const int wid = blockIdx.y * VERIFY_WARPS + (threadIdx.x >> 5);
  const int nwarps = gridDim.y * VERIFY_WARPS;
  const Cand* src = cands + p.cand_base;
  Rec* dst = out + offsets[blockIdx.x];
  if (p.mode != PM_TC) {      // exact / Hamming pairs, and tensor-core pairs whose candidates were already re-scored in-kernel
    for (int c = wid * 32 + lane; c < n; c += nwarps * 32) {
      const Cand k = src[c];
      dst[c] = Rec{k.b, k.q, k.d1, k.d2};
    }
    return;
  }
  const ViewDev& vi = views[p.view_i];
  const ViewDev& vj = views[p.view_j];
  for (int c = wid; c < n; c += nwarps) {
    const Cand k = src[c];
    Rec rec;
    const bool keep = rescore_candidate(vi.h16, vj.h16, p.m_i, k, ratio_sq, lane, err_count, rec);
    if (lane == 0) dst[c] = keep ? rec : Rec{0xFFFFFFFFu, k.q, rec.d1, rec.d2};
  }
}

}  // namespace b200m
