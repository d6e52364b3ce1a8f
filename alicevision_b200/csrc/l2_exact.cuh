// Exact CUDA-core search kernels (no tensor cores): the path for descriptors the tensor-core kernel cannot
// take bit-exactly (real-valued fp32, out-of-range values, dim != 128) and for the ArrayMatcher surface.
//
//  * exact_top2_kernel<T>: dim == 128, two nearest neighbours, 64 queries x 64 database rows per smem tile.
//    float  -> the reference's SSE summation order, feature/metric.hpp:94-123: four independent lanes
//              s_l += (a-b)*(a-b) (separate multiply and add, no FMA), result ((s0+s1)+s2)+s3  => bit-exact.
//    uint8  -> feature/metric.hpp:48-80; all terms are exact integers < 2^24 so any order is bit-exact.
//  * generic_knn_kernel: any dim, NN <= 16, metric L2_Simple (sequential order, metric.hpp:27-44),
//    L2_Vectorized or Hamming (feature/Hamming.hpp:113-148); one warp per query.
// Ties are ordered by (distance, database index); the reference leaves tie order unspecified
// (matching/matching_test.cpp:46).
#pragma once
#include "common.cuh"

namespace b200m {

struct T2 { float m1, m2; int i1, i2; };   // two smallest (value, index), lexicographic

__device__ __forceinline__ bool lt(float a, int ia, float b, int ib) { return a < b || (a == b && ia < ib); }
__device__ __forceinline__ void t2_push(T2& s, float v, int i) {
  if (lt(v, i, s.m1, s.i1)) { s.m2 = s.m1; s.i2 = s.i1; s.m1 = v; s.i1 = i; }
  else if (lt(v, i, s.m2, s.i2)) { s.m2 = v; s.i2 = i; }
}
__device__ __forceinline__ void t2_merge(T2& s, const T2& o) { t2_push(s, o.m1, o.i1); t2_push(s, o.m2, o.i2); }

constexpr int EX_TQ = 64, EX_TD = 64, EX_LD = 132;   // tile sizes, padded smem row stride (floats)
constexpr int EX_SMEM = (EX_TQ + EX_TD) * EX_LD * 4;

// grid = (ceil(m_j / 64), n_pairs); pairs[p].mode selects float / uint8 interpretation via template dispatch on host.
// OUT_DENSE: write (idx, dist) for both neighbours to dense arrays (ArrayMatcher surface);
// otherwise apply the ratio test and append a candidate (collection surface).
template <typename T, bool OUT_DENSE>
__global__ void __launch_bounds__(256)
exact_top2_kernel(const ViewDev* __restrict__ views, const PairDev* __restrict__ pairs, uint32_t want_mode, Cand* __restrict__ cands,
                  int* __restrict__ cand_count, float ratio_sq, int32_t* __restrict__ dense_idx, float* __restrict__ dense_dist) {
  const PairDev p = pairs[blockIdx.y];
  if (p.mode != want_mode) return;
  const int q0 = blockIdx.x * EX_TQ;
  if (q0 >= (int)p.m_j) return;
  extern __shared__ float exsm[];
  float* Qs = exsm;
  float* Ds = exsm + EX_TQ * EX_LD;
  // T names the arithmetic (float: SSE order; uint8: exact integers - identical results whenever both apply); the element type of
  // each view's buffer is its own storage type (integer-valued fp32 views are stored as u8)
  const ViewDev& vq = views[p.view_j];
  const ViewDev& vd = views[p.view_i];
  const int tid = threadIdx.x, ty = tid >> 4, tx = tid & 15;

  for (int e = tid; e < EX_TQ * 128; e += 256) {
    const int r = e >> 7, c = e & 127;
    Qs[r * EX_LD + c] = (q0 + r < (int)p.m_j) ? view_elem(vq, (size_t)(q0 + r) * 128 + c) : 0.f;
  }
  T2 best[4];
#pragma unroll
  for (int i = 0; i < 4; ++i) best[i] = T2{INFINITY, INFINITY, 0x7fffffff, 0x7fffffff};

  for (int d0 = 0; d0 < (int)p.m_i; d0 += EX_TD) {
    __syncthreads();
    for (int e = tid; e < EX_TD * 128; e += 256) {
      const int r = e >> 7, c = e & 127;
      Ds[r * EX_LD + c] = (d0 + r < (int)p.m_i) ? view_elem(vd, (size_t)(d0 + r) * 128 + c) : 0.f;
    }
    __syncthreads();
    float acc[4][4][4];
#pragma unroll
    for (int a = 0; a < 4; ++a)
#pragma unroll
      for (int b = 0; b < 4; ++b)
#pragma unroll
        for (int l = 0; l < 4; ++l) acc[a][b][l] = 0.f;
#pragma unroll 2
    for (int k = 0; k < 128; k += 4) {
      float4 qa[4], db[4];
#pragma unroll
      for (int a = 0; a < 4; ++a) qa[a] = *reinterpret_cast<const float4*>(&Qs[(ty * 4 + a) * EX_LD + k]);
#pragma unroll
      for (int b = 0; b < 4; ++b) db[b] = *reinterpret_cast<const float4*>(&Ds[(tx + 16 * b) * EX_LD + k]);
#pragma unroll
      for (int a = 0; a < 4; ++a)
#pragma unroll
        for (int b = 0; b < 4; ++b) {
          float t;
          t = __fsub_rn(qa[a].x, db[b].x); acc[a][b][0] = __fadd_rn(acc[a][b][0], __fmul_rn(t, t));
          t = __fsub_rn(qa[a].y, db[b].y); acc[a][b][1] = __fadd_rn(acc[a][b][1], __fmul_rn(t, t));
          t = __fsub_rn(qa[a].z, db[b].z); acc[a][b][2] = __fadd_rn(acc[a][b][2], __fmul_rn(t, t));
          t = __fsub_rn(qa[a].w, db[b].w); acc[a][b][3] = __fadd_rn(acc[a][b][3], __fmul_rn(t, t));
        }
    }
#pragma unroll
    for (int b = 0; b < 4; ++b) {
      const int di = d0 + tx + 16 * b;
      if (di < (int)p.m_i) {
#pragma unroll
        for (int a = 0; a < 4; ++a) {
          const float d = __fadd_rn(__fadd_rn(__fadd_rn(acc[a][b][0], acc[a][b][1]), acc[a][b][2]), acc[a][b][3]);
          t2_push(best[a], d, di);
        }
      }
    }
  }
  // merge the 16 threads (tx) that share a query: they are 16 consecutive lanes
#pragma unroll
  for (int a = 0; a < 4; ++a) {
#pragma unroll
    for (int o = 8; o >= 1; o >>= 1) {
      T2 oth;
      oth.m1 = __shfl_xor_sync(0xffffffffu, best[a].m1, o); oth.i1 = __shfl_xor_sync(0xffffffffu, best[a].i1, o);
      oth.m2 = __shfl_xor_sync(0xffffffffu, best[a].m2, o); oth.i2 = __shfl_xor_sync(0xffffffffu, best[a].i2, o);
      t2_merge(best[a], oth);
    }
    const int q = q0 + ty * 4 + a;
    if (tx == 0 && q < (int)p.m_j) {
      if (OUT_DENSE) {
        dense_idx[2 * q] = best[a].i1; dense_idx[2 * q + 1] = best[a].i2;
        dense_dist[2 * q] = best[a].m1; dense_dist[2 * q + 1] = best[a].m2;
      } else if (best[a].m1 < __fmul_rn(ratio_sq, best[a].m2)) {      // matching/filters.hpp:60
        const int slot = atomicAdd(&cand_count[blockIdx.y], 1);
        cands[p.cand_base + slot] = Cand{(uint32_t)q, (uint32_t)best[a].i1, best[a].m1, best[a].m2};
      }
    }
  }
}

// ---------------------------------------------------------------------------------------------------------------
// Generic k-NN: one warp per query, lanes stride over database rows, per-lane sorted list of the NN best.
enum : int { MET_L2_SIMPLE = 0, MET_L2_VECTORIZED = 1, MET_HAMMING = 2 };
constexpr int GEN_MAX_NN = 16;

template <typename T>
__device__ __forceinline__ float gen_dist(const T* __restrict__ q, const T* __restrict__ d, int dim, int metric) {
  if (metric == MET_L2_VECTORIZED && sizeof(T) == 4) {
    if (dim & 3) return 0.f;                       // metric.hpp:118-122: "size is not modulus 4" -> 0
    float s0 = 0.f, s1 = 0.f, s2 = 0.f, s3 = 0.f;
    for (int k = 0; k < dim; k += 4) {
      float t;
      t = __fsub_rn((float)q[k], (float)d[k]); s0 = __fadd_rn(s0, __fmul_rn(t, t));
      t = __fsub_rn((float)q[k + 1], (float)d[k + 1]); s1 = __fadd_rn(s1, __fmul_rn(t, t));
      t = __fsub_rn((float)q[k + 2], (float)d[k + 2]); s2 = __fadd_rn(s2, __fmul_rn(t, t));
      t = __fsub_rn((float)q[k + 3], (float)d[k + 3]); s3 = __fadd_rn(s3, __fmul_rn(t, t));
    }
    return __fadd_rn(__fadd_rn(__fadd_rn(s0, s1), s2), s3);
  }
  if (metric == MET_L2_VECTORIZED) {               // non-float: metric.hpp:48-80, r += d0^2+d1^2+d2^2+d3^2 per group of 4
    float r = 0.f; int k = 0;
    for (; k + 3 < dim; k += 4) {
      const float d0 = (float)q[k] - (float)d[k], d1 = (float)q[k + 1] - (float)d[k + 1];
      const float d2 = (float)q[k + 2] - (float)d[k + 2], d3 = (float)q[k + 3] - (float)d[k + 3];
      r = __fadd_rn(r, __fadd_rn(__fadd_rn(__fadd_rn(__fmul_rn(d0, d0), __fmul_rn(d1, d1)), __fmul_rn(d2, d2)), __fmul_rn(d3, d3)));
    }
    for (; k < dim; ++k) { const float t = (float)q[k] - (float)d[k]; r = __fadd_rn(r, __fmul_rn(t, t)); }
    return r;
  }
  float r = 0.f;                                    // L2_Simple, metric.hpp:27-44
  for (int k = 0; k < dim; ++k) { const float t = __fsub_rn((float)q[k], (float)d[k]); r = __fadd_rn(r, __fmul_rn(t, t)); }
  return r;
}
__device__ __forceinline__ float gen_hamming(const uint8_t* __restrict__ q, const uint8_t* __restrict__ d, int nbytes) {
  unsigned r = 0; int k = 0;
  if (((reinterpret_cast<uintptr_t>(q) | reinterpret_cast<uintptr_t>(d)) & 3) == 0)
    for (; k + 3 < nbytes; k += 4) r += __popc(*reinterpret_cast<const uint32_t*>(q + k) ^ *reinterpret_cast<const uint32_t*>(d + k));
  for (; k < nbytes; ++k) r += __popc((unsigned)(q[k] ^ d[k]));
  return __uint_as_float(r);                       // carried as raw bits; compared as unsigned below
}

// dist_bits: float bit patterns for L2 (non-negative floats order like unsigned ints) or uint32 Hamming distances.
template <typename T>
__global__ void __launch_bounds__(128)
generic_knn_kernel(const T* __restrict__ db, int n_db, const T* __restrict__ qs, int n_q, int dim, int nn, int metric,
                   int32_t* __restrict__ out_idx, uint32_t* __restrict__ out_dist_bits) {
  const int q = blockIdx.x * 4 + (threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
  if (q >= n_q) return;
  const T* qp = qs + (size_t)q * dim;
  uint32_t bv[GEN_MAX_NN]; int bi[GEN_MAX_NN];
#pragma unroll
  for (int k = 0; k < GEN_MAX_NN; ++k) { bv[k] = 0xFFFFFFFFu; bi[k] = 0x7fffffff; }
  for (int r = lane; r < n_db; r += 32) {
    const T* dp = db + (size_t)r * dim;
    float dv;
    if (metric == MET_HAMMING) dv = gen_hamming(reinterpret_cast<const uint8_t*>(qp), reinterpret_cast<const uint8_t*>(dp), dim);
    else dv = gen_dist(qp, dp, dim, metric);
    uint32_t v = __float_as_uint(dv); int idx = r;
#pragma unroll
    for (int k = 0; k < GEN_MAX_NN; ++k) {          // sorted insertion by (value, index)
      if (k < nn && (v < bv[k] || (v == bv[k] && idx < bi[k]))) {
        const uint32_t tv = bv[k]; const int ti = bi[k]; bv[k] = v; bi[k] = idx; v = tv; idx = ti;
      }
    }
  }
  // warp merge: nn rounds of "take the global minimum head"
  for (int k = 0; k < nn; ++k) {
    uint32_t hv = bv[0]; int hi = bi[0]; int src = lane;
#pragma unroll
    for (int o = 16; o >= 1; o >>= 1) {
      const uint32_t ov = __shfl_xor_sync(0xffffffffu, hv, o); const int oi = __shfl_xor_sync(0xffffffffu, hi, o);
      const int os = __shfl_xor_sync(0xffffffffu, src, o);
      if (ov < hv || (ov == hv && oi < hi)) { hv = ov; hi = oi; src = os; }
    }
    if (lane == 0) { out_idx[(size_t)q * nn + k] = hi; out_dist_bits[(size_t)q * nn + k] = hv; }
    if (lane == src) {
#pragma unroll
      for (int j = 0; j + 1 < GEN_MAX_NN; ++j) { bv[j] = bv[j + 1]; bi[j] = bi[j + 1]; }
      bv[GEN_MAX_NN - 1] = 0xFFFFFFFFu; bi[GEN_MAX_NN - 1] = 0x7fffffff;
    }
  }
}

// Collection surface for scalar descriptors of any length other than 128 (AKAZE_Float_Regions = float x 64, AKAZE_Liop_Regions =
// uchar x 144; feature/regionsFactory.hpp:25-27): one warp per query of pair blockIdx.y, L2_Vectorized in the reference's order
// (gen_dist), two nearest neighbours by (value, index), ratio test, candidate appended.  A functional path, not a tuned one.
template <typename T>
__global__ void __launch_bounds__(128)
generic_top2_pairs_kernel(const ViewDev* __restrict__ views, const PairDev* __restrict__ pairs, uint32_t want_mode, Cand* __restrict__ cands,
                          int* __restrict__ cand_count, float ratio_sq) {
  const PairDev p = pairs[blockIdx.y];
  if (p.mode != want_mode) return;
  const int q = blockIdx.x * 4 + (threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
  if (q >= (int)p.m_j) return;
  const int dim = views[p.view_i].dim;
  const T* db = reinterpret_cast<const T*>(views[p.view_i].raw);
  const T* qp = reinterpret_cast<const T*>(views[p.view_j].raw) + (size_t)q * dim;
  uint32_t v1 = 0xFFFFFFFFu, v2 = 0xFFFFFFFFu; int i1 = 0x7fffffff, i2 = 0x7fffffff;
  for (int r = lane; r < (int)p.m_i; r += 32) {
    const uint32_t v = __float_as_uint(gen_dist(qp, db + (size_t)r * dim, dim, MET_L2_VECTORIZED));   // non-negative floats order like their bit patterns
    if (v < v1 || (v == v1 && r < i1)) { v2 = v1; i2 = i1; v1 = v; i1 = r; }
    else if (v < v2 || (v == v2 && r < i2)) { v2 = v; i2 = r; }
  }
#pragma unroll
  for (int o = 16; o >= 1; o >>= 1) {                        // merge two sorted pairs per step
    const uint32_t a1 = __shfl_xor_sync(0xffffffffu, v1, o), a2 = __shfl_xor_sync(0xffffffffu, v2, o);
    const int j1 = __shfl_xor_sync(0xffffffffu, i1, o), j2 = __shfl_xor_sync(0xffffffffu, i2, o);
    if (a1 < v1 || (a1 == v1 && j1 < i1)) {
      if (v1 < a2 || (v1 == a2 && i1 < j2)) { v2 = v1; i2 = i1; } else { v2 = a2; i2 = j2; }
      v1 = a1; i1 = j1;
    } else if (a1 < v2 || (a1 == v2 && j1 < i2)) { v2 = a1; i2 = j1; }
  }
  if (lane == 0) {
    const float d1 = __uint_as_float(v1), d2 = __uint_as_float(v2);
    if (d1 < __fmul_rn(ratio_sq, d2)) {                      // matching/filters.hpp:60
      const int slot = atomicAdd(&cand_count[blockIdx.y], 1);
      cands[p.cand_base + slot] = Cand{(uint32_t)q, (uint32_t)i1, d1, d2};
    }
  }
}

}  // namespace b200m
