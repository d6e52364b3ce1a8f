// Device-side finishing of RegionsMatcher::Match's tail (matching/RegionsMatcher.hpp:153-175) for views in general position.
//
// The reference turns the ratio-test survivors of a directed pair into the final IndMatches by
//   1. IndMatch::getDeduplicated                (matching/IndMatch.hpp:52-58): std::set ordered by (i, j) - every query j occurs once,
//      so this is only a sort by (i, j);
//   2. IndMatchDecorator::getDeduplicated       (matching/IndMatchDecorator.hpp:57-69,84-98): std::set over (x_i, y_i, x_j, y_j) with a
//      comparator that is not a strict weak order.  When all left x are distinct and all left y are distinct ("general position",
//      checked per view by pos_rank_kernel) its effect is exactly: keep the FIRST inserted match of every left feature i (= the
//      smallest j), and list the kept matches by ascending left y.
// Both steps are index arithmetic, so for such views the whole tail runs on the GPU (round-1 verdict: the host finishing stage was
// the wall of the end-to-end number at >= 4 ranks): min-j per left feature by atomicMin, then a direct-address scatter by the
// precomputed rank of y_i and an ordered compaction.  Views with colliding coordinates keep the literal host std::set path.
#pragma once
#include "common.cuh"

namespace b200m {

// yrank[i] = number of features of the view whose y is smaller than y_i (the position of feature i in the ascending-y order when all
// y are distinct); sets VF_POS_NONGENERIC in *flags when two features share an x or a y, or a coordinate is NaN.
// O(m^2) comparisons spread over grid (i blocks) x (k ranges): partial counts are added atomically into yrank (zeroed by the
// caller).  No shared memory on purpose: these launches run on the upload stream WHILE the persistent tensor-core kernel owns
// every SM with ~193 KB of shared memory per CTA - a block that needs its own shared-memory carve-out would have to wait for that
// kernel to end (measured: 0.65 ms per view of upload stall); the positions are read as warp-uniform (broadcast) loads instead.
constexpr int PR_THREADS = 128, PR_KRANGE = 1024;
__global__ void __launch_bounds__(PR_THREADS)
pos_rank_kernel(const float2* __restrict__ xy, int m, uint32_t* __restrict__ yrank, uint32_t* __restrict__ flags) {
  const int i = blockIdx.x * PR_THREADS + threadIdx.x;
  const int k0 = blockIdx.y * PR_KRANGE, k1 = min(k0 + PR_KRANGE, m);
  const float2 me = i < m ? xy[i] : make_float2(0.f, 0.f);
  uint32_t less = 0; bool dup = false;
#pragma unroll 8
  for (int k = k0; k < k1; ++k) {
    const float2 o = __ldg(xy + k);                      // same address on every lane: one broadcast transaction, L1-resident
    less += (o.y < me.y) ? 1u : 0u;
    dup |= (k != i) && (o.x == me.x || o.y == me.y);
  }
  if (i < m) {
    if (gridDim.y == 1) yrank[i] = less;                 // one k range: the count is complete (the caller did not zero yrank)
    else if (less) atomicAdd(&yrank[i], less);
    if (dup || (blockIdx.y == 0 && (me.x != me.x || me.y != me.y))) atomicOr(flags, VF_POS_NONGENERIC);
  }
}

// One block per directed pair of the batch.  recs = packed ratio-test survivors (Rec, i == 0xFFFFFFFF = dropped by the exactness
// pass) at offsets[p]..+count[p]; fin receives, at the same offsets, either the FINAL IndMatches (fin_count[p] >= 0 of them:
// i, j, distance ratio, distance - matching/IndMatch.hpp:25-65, RegionsMatcher.hpp:157-158) or, for a database view that is not
// in general position, the records unchanged with fin_count[p] = -(count + 1) so that the host finishes that pair.
// scratch: 2 x m_i uint32 per pair at 2 * slot_base.
constexpr int FIN_THREADS = 512;
struct FinMatch { uint32_t i, j; float ratio, dist; };

__global__ void __launch_bounds__(FIN_THREADS)
finish_pairs_kernel(const ViewDev* __restrict__ views, const PairDev* __restrict__ pairs, const uint32_t* __restrict__ view_flags,
                    const Rec* __restrict__ recs, const int* __restrict__ count, const int* __restrict__ offsets,
                    uint32_t* __restrict__ scratch, FinMatch* __restrict__ fin, int* __restrict__ fin_count) {
  const PairDev p = pairs[blockIdx.x];
  const int n = count[blockIdx.x];
  const int tid = threadIdx.x;
  if (n == 0 || p.mode == PM_SKIP) { if (tid == 0) fin_count[blockIdx.x] = 0; return; }
  const Rec* src = recs + offsets[blockIdx.x];
  FinMatch* dst = fin + offsets[blockIdx.x];
  const ViewDev& vi = views[p.view_i];
  const uint32_t* yrank = vi.yrank;
  if (yrank == nullptr || (view_flags[p.view_i] & VF_POS_NONGENERIC)) {
    for (int k = tid; k < n; k += FIN_THREADS) { const Rec r = src[k]; dst[k] = FinMatch{r.i, r.j, r.d1, r.d2}; }
    if (tid == 0) fin_count[blockIdx.x] = -(n + 1);
    return;
  }
  const int m_i = (int)p.m_i;
  uint32_t* A = scratch + 2 * (size_t)p.slot_base;    // min query j per database feature i
  uint32_t* B = A + m_i;                              // record index by rank of y_i
  for (int k = tid; k < 2 * m_i; k += FIN_THREADS) A[k] = 0xFFFFFFFFu;
  __syncthreads();
  for (int k = tid; k < n; k += FIN_THREADS) {
    const Rec r = src[k];
    if (r.i < (uint32_t)m_i) atomicMin(&A[r.i], r.j);                 // dropped records carry i = 0xFFFFFFFF
  }
  __syncthreads();
  for (int k = tid; k < n; k += FIN_THREADS) {
    const Rec r = src[k];
    if (r.i < (uint32_t)m_i && A[r.i] == r.j) B[yrank[r.i]] = (uint32_t)k;   // ranks are distinct in general position
  }
  __syncthreads();
  // ordered compaction of B: thread t owns the contiguous segment [t*S, (t+1)*S)
  const int S = (m_i + FIN_THREADS - 1) / FIN_THREADS;
  const int s0 = min(tid * S, m_i), s1 = min(s0 + S, m_i);
  int mine = 0;
  for (int k = s0; k < s1; ++k) mine += (B[k] != 0xFFFFFFFFu);
  __shared__ int wsum[FIN_THREADS / 32];
  __shared__ int total;
  const int lane = tid & 31, warp = tid >> 5;
  int x = mine;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) { const int y = __shfl_up_sync(0xffffffffu, x, o); if (lane >= o) x += y; }
  if (lane == 31) wsum[warp] = x;
  __syncthreads();
  if (warp == 0) {
    int s = lane < FIN_THREADS / 32 ? wsum[lane] : 0;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) { const int y = __shfl_up_sync(0xffffffffu, s, o); if (lane >= o) s += y; }
    if (lane < FIN_THREADS / 32) wsum[lane] = s;
    if (lane == FIN_THREADS / 32 - 1) total = s;
  }
  __syncthreads();
  int w = (warp ? wsum[warp - 1] : 0) + x - mine;
  const bool hamming = p.mode == PM_HAMMING;
  for (int k = s0; k < s1; ++k) {
    const uint32_t e = B[k];
    if (e == 0xFFFFFFFFu) continue;
    const Rec r = src[e];
    FinMatch f;
    f.i = r.i; f.j = r.j;
    if (hamming) {
      const uint32_t d1 = __float_as_uint(r.d1), d2 = __float_as_uint(r.d2);
      f.ratio = (float)(d1 / d2);                     // integer division, matching/filters.hpp:64 on unsigned (d2 > 0 for a survivor)
      f.dist = (float)d1;
    } else {
      f.ratio = __fdiv_rn(r.d1, r.d2);                // float division, filters.hpp:64
      f.dist = r.d1;
    }
    dst[w++] = f;
  }
  if (tid == 0) fin_count[blockIdx.x] = total;
}

// u8 -> f32 expansion of a view that was staged as uchar (integer-valued fp32 descriptors) for the rare consumer that needs both
// views of a pair in one element type.
__global__ void u8_to_f32_kernel(const uint8_t* __restrict__ src, float* __restrict__ dst, size_t n) {
  const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) dst[i] = (float)src[i];
}

}  // namespace b200m
