// K2 — per-view preparation at upload time (runs once per image, not on the pair hot path):
// fp16 copy of the 128-D descriptors for the tensor-core kernel, squared row norms, and the exactness
// flags that decide whether a view may take the tensor-core path (see common.cuh VF_*).
#pragma once
#include "common.cuh"

namespace b200m {

// One warp per row; lane handles components lane, lane+32, lane+64, lane+96.
template <typename T>
__global__ void __launch_bounds__(256)
prep_view_kernel(const T* __restrict__ raw, int m, __half* __restrict__ h16, float* __restrict__ nbh, float* __restrict__ nrm,
                 int m_pad, uint32_t* __restrict__ flags, __half* __restrict__ aug16) {
  const int row = blockIdx.x * 8 + (threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
  if (row >= m_pad) return;
  if (row >= m) {     // pad rows: a half-norm larger than any real one (2048 * 2047 = 2^22 - 2048)
    if (lane == 0) nbh[row] = 1e30f;
    if (lane < 16) aug16[(size_t)row * 16 + lane] = __float2half_rn(lane == 2 ? 2047.f : 0.f);
    return;
  }
  float s = 0.f; uint32_t f = 0;
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    const float v = (float)raw[(size_t)row * 128 + lane + 32 * k];
    if (v != rintf(v)) f |= VF_NONINTEGER;
    if (!(fabsf(v) <= 1024.f)) f |= VF_RANGE;
    h16[(size_t)row * 128 + lane + 32 * k] = __float2half_rn(v);
    s = fmaf(v, v, s);
  }
#pragma unroll
  for (int o = 16; o >= 1; o >>= 1) { s += __shfl_xor_sync(0xffffffffu, s, o); f |= __shfl_xor_sync(0xffffffffu, f, o); }
  if (lane == 0) {
    if (!(s < 4194304.f)) f |= VF_NORM;
    nrm[row] = s;
    nbh[row] = 0.5f * s;
    if (f) atomicOr(flags, f);
  }
  // limbs of ||row||^2 / 2 (exact when the norm is an integer < 2^22, i.e. whenever the view is tensor-core eligible)
  const float b0 = fmodf(s, 2.f), rest = (s - b0) * 0.5f, l1 = floorf(rest * (1.f / 2048.f)), l0 = rest - 2048.f * l1;
  if (lane < 16) aug16[(size_t)row * 16 + lane] = __float2half_rn(lane == 0 ? b0 : lane == 1 ? l0 : lane == 2 ? fminf(l1, 2047.f) : 0.f);
}

}  // namespace b200m
