// K7 — guided matching: descriptor top-2 + ratio test restricted to the correspondences that agree with a fundamental
// matrix.  Replaces matching::guidedMatching<Mat3Model, FundamentalEpipolarDistanceError>(model, camL,
// lRegions, camR, rRegions, errorTh, distRatio, matches) for cameras without distortion (matching/guidedMatching.hpp:206-268)
// with the error of multiview/relativePose/FundamentalError.hpp:52-64 and the accumulator matching/guidedMatching.hpp:77-118.
//
// The geometric predicate passes ~0.1-1 % of the (left, right) pairs (a band of +-sqrt(errorTh) pixels around the epipolar
// line), so this is not tensor-core work: the O(M_l x M_r) part is ~12 double-precision operations per pair (done exactly as
// the reference does them: separate multiply / add / divide, no contraction), and only the survivors pay for a 128-D
// descriptor distance, computed with the reference's own summation order so that distances, hence the ratio test, are
// bit-identical (float: feature/metric.hpp:94-123 four sequential lanes, ((s0+s1)+s2)+s3; uchar: exact integers;
// binary: feature/Hamming.hpp:174-185 squared popcount).
//
// One warp per left feature; the right positions stream through shared memory; survivors are queued per warp and their
// distances evaluated eight at a time (4 lanes per candidate = the 4 SSE lanes).
#pragma once
#include "common.cuh"

#include <cfloat>

namespace b200m {

constexpr int GM_WARPS = 8;          // left features per block
constexpr int GM_TILE = 1024;        // right positions per shared-memory tile
constexpr int GM_QUEUE = 40;         // >= 8 + 32: survivors of one 32-wide step appended to at most 7 waiting ones

enum : int { GM_FUNDAMENTAL = 0, GM_HOMOGRAPHY = 1 };

struct GuidedParams {
  int model;            // GM_FUNDAMENTAL: FundamentalEpipolarDistanceError; GM_HOMOGRAPHY: HomographyAsymmetricError (HomographyError.hpp:23-31)
  double F[9];          // row-major model matrix: fundamental (x_right^T F x_left = 0) or homography (x_right ~ H x_left)
  double errorTh;       // Square(precision), guidedMatching.hpp:211 / GeometricFilterMatrix_F_AC.hpp:387
  double distRatio;     // Square(distance ratio), :388
};

// distance of candidate descriptors a (left) and b (right) evaluated by the 4 lanes of a group; returns the value on every lane of the group
template <int DTYPE>
__device__ __forceinline__ double guided_distance(const void* __restrict__ left, const void* __restrict__ right, int i, int j, int sub, unsigned gmask) {
  if (DTYPE == DT_BIN) {
    const uint32_t* a = reinterpret_cast<const uint32_t*>(left) + (size_t)i * 16;
    const uint32_t* b = reinterpret_cast<const uint32_t*>(right) + (size_t)j * 16;
    unsigned h = 0;
#pragma unroll
    for (int w = 0; w < 4; ++w) h += __popc(a[sub * 4 + w] ^ b[sub * 4 + w]);
    h += __shfl_xor_sync(gmask, h, 1);
    h += __shfl_xor_sync(gmask, h, 2);
    return (double)(h * h);                                           // SquaredHamming: h * h in unsigned, then double
  }
  float s = 0.f;
  if (DTYPE == DT_F32) {
    const float* a = reinterpret_cast<const float*>(left) + (size_t)i * 128;
    const float* b = reinterpret_cast<const float*>(right) + (size_t)j * 128;
    for (int t = 0; t < 32; ++t) {                                    // SSE lane `sub`: s += (a-b)*(a-b), t ascending (metric.hpp:100-110)
      const float d = __fsub_rn(a[4 * t + sub], b[4 * t + sub]);
      s = __fadd_rn(s, __fmul_rn(d, d));
    }
  } else {
    const uint8_t* a = reinterpret_cast<const uint8_t*>(left) + (size_t)i * 128;
    const uint8_t* b = reinterpret_cast<const uint8_t*>(right) + (size_t)j * 128;
    int acc = 0;
    for (int t = 0; t < 32; ++t) { const int d = (int)a[4 * t + sub] - (int)b[4 * t + sub]; acc += d * d; }
    s = (float)acc;                                                   // exact integers < 2^24: any order gives the reference's float
  }
  const float s1 = __shfl_xor_sync(gmask, s, 1);                      // lanes (0,1) and (2,3)
  const float p = (sub & 1) ? __fadd_rn(s1, s) : __fadd_rn(s, s1);    // s0+s1 on lanes 0,1 ; s2+s3 on lanes 2,3
  const float s0123 = __shfl_sync(gmask, p, 0, 4);                    // (s0+s1)
  const float s2 = __shfl_sync(gmask, s, 2, 4), s3 = __shfl_sync(gmask, s, 3, 4);
  return (double)__fadd_rn(__fadd_rn(s0123, s2), s3);                 // ((s0+s1)+s2)+s3, metric.hpp:112-116
}

template <int DTYPE>
__global__ void __launch_bounds__(GM_WARPS * 32)
guided_top2_kernel(const void* __restrict__ left, const void* __restrict__ right, const double2* __restrict__ xl, const double2* __restrict__ xr,
                   int m_l, int m_r, GuidedParams P, Rec* __restrict__ out, int* __restrict__ out_count) {
  __shared__ double2 tile[GM_TILE];
  __shared__ int queue[GM_WARPS][GM_QUEUE];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int i = blockIdx.x * GM_WARPS + warp;
  const bool live = i < m_l;
  // epipolar line of the left point in the right image: F_x = F * (x, y, 1), row r = (F(r,0)*x + F(r,1)*y) + F(r,2)*1
  double fx0 = 0, fx1 = 0, fx2 = 0, nrm = 1;
  if (live) {
    const double2 x = xl[i];
    fx0 = __dadd_rn(__dadd_rn(__dmul_rn(P.F[0], x.x), __dmul_rn(P.F[1], x.y)), P.F[2]);
    fx1 = __dadd_rn(__dadd_rn(__dmul_rn(P.F[3], x.x), __dmul_rn(P.F[4], x.y)), P.F[5]);
    fx2 = __dadd_rn(__dadd_rn(__dmul_rn(P.F[6], x.x), __dmul_rn(P.F[7], x.y)), P.F[8]);
    nrm = __dadd_rn(__dmul_rn(fx0, fx0), __dmul_rn(fx1, fx1));      // F_x.head<2>().squaredNorm()
    if (P.model == GM_HOMOGRAPHY) { fx0 = __ddiv_rn(fx0, fx2); fx1 = __ddiv_rn(fx1, fx2); }   // x2_est = x2h_est.head<2>() / x2h_est[2]
  }
  double bd = DBL_MAX, sbd = DBL_MAX; int idx = 0;                   // distanceRatio<double>, guidedMatching.hpp:77-90
  int nq = 0;
  const int sub = lane & 3, grp = lane >> 2;
  const unsigned gmask = 0xFu << (grp * 4);

  auto drain = [&](int count) {                                       // evaluate `count` (<= 8) queued candidates, oldest first
    const int j = grp < count ? queue[warp][grp] : -1;
    double d = 0;
    if (j >= 0) d = guided_distance<DTYPE>(left, right, i, j, sub, gmask);
    for (int g = 0; g < count; ++g) {                                 // distanceRatio::update in ascending j (:95-110)
      const double dist = __shfl_sync(0xffffffffu, d, g * 4);
      const int jj = __shfl_sync(0xffffffffu, j, g * 4);
      if (dist < bd) { idx = jj; sbd = bd; bd = dist; }
      else if (dist < sbd) sbd = dist;
    }
  };

  for (int t0 = 0; t0 < m_r; t0 += GM_TILE) {
    __syncthreads();
    for (int e = threadIdx.x; e < GM_TILE && t0 + e < m_r; e += GM_WARPS * 32) tile[e] = xr[t0 + e];
    __syncthreads();
    if (!live) continue;
    const int n = min(GM_TILE, m_r - t0);
    for (int b = 0; b < n; b += 32) {
      const int e = b + lane;
      bool pass = false;
      if (e < n) {
        const double2 y = tile[e];
        double err;
        if (P.model == GM_HOMOGRAPHY) {
          const double d0 = __dsub_rn(y.x, fx0), d1 = __dsub_rn(y.y, fx1);                            // (x2 - x2_est).squaredNorm()
          err = __dadd_rn(__dmul_rn(d0, d0), __dmul_rn(d1, d1));
        } else {
          const double dot = __dadd_rn(__dadd_rn(__dmul_rn(fx0, y.x), __dmul_rn(fx1, y.y)), fx2);    // F_x.dot((y, 1))
          err = __ddiv_rn(__dmul_rn(dot, dot), nrm);                                                  // Square(dot) / squaredNorm
        }
        pass = err < P.errorTh;                                                                       // guidedMatching.hpp:252
      }
      const unsigned m = __ballot_sync(0xffffffffu, pass);
      if (m) {
        if (pass) queue[warp][nq + __popc(m & ((1u << lane) - 1))] = t0 + e;
        nq += __popc(m);
        __syncwarp();
        while (nq >= 8) {
          drain(8);
          __syncwarp();
          const int rest = nq - 8;                                    // shift the remaining entries to the front
          const int v = lane < rest ? queue[warp][8 + lane] : 0;
          __syncwarp();
          if (lane < rest) queue[warp][lane] = v;
          nq = rest;
          __syncwarp();
        }
      }
    }
  }
  if (!live) return;
  if (nq > 0) drain(nq);
  // distanceRatio::isValid (:115-118) and the emitted pair (:259-263)
  if (lane == 0 && sbd != DBL_MAX && bd < __dmul_rn(P.distRatio, sbd)) {
    const int slot = atomicAdd(out_count, 1);
    out[slot] = Rec{(uint32_t)i, (uint32_t)idx, (float)bd, (float)sbd};
  }
}

__global__ void positions_to_double_kernel(const float* __restrict__ xy, int n, double2* __restrict__ out) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) out[i] = make_double2((double)xy[2 * i], (double)xy[2 * i + 1]);
}

}  // namespace b200m
