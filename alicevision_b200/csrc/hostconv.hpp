// Host-side staging helper: integer-valued fp32 descriptors (what the reference's SIFT extractor stores in SIFT_Float_Regions:
// floor(512*sqrt(x)), feature/sift/SIFT.hpp:80-110) are converted to uchar while they are copied into the pinned staging ring,
// so 4x fewer bytes cross PCIe and the host writes a quarter of the bytes.  The conversion is CHECKED: the first element
// that is not exactly representable as a uchar aborts it and the view is uploaded as fp32 instead.
#pragma once
#include <cstddef>
#include <cstdint>
#if defined(__x86_64__)
#include <immintrin.h>
#endif

namespace b200m {

inline bool f32_to_u8_checked_scalar(const float* src, uint8_t* dst, size_t n) {
  for (size_t k = 0; k < n; ++k) {
    const float v = src[k];
    if (!(v >= 0.f && v <= 255.f)) return false;
    const uint8_t u = (uint8_t)v;
    if ((float)u != v) return false;
    dst[k] = u;
  }
  return true;
}

#if defined(__x86_64__)
__attribute__((target("avx2"))) inline bool f32_to_u8_checked_avx2(const float* src, uint8_t* dst, size_t n) {
  size_t k = 0;
  const __m256i perm = _mm256_setr_epi32(0, 4, 1, 5, 2, 6, 3, 7);
  __m256 bad = _mm256_setzero_ps();
  for (; k + 32 <= n; k += 32) {
    const __m256 a = _mm256_loadu_ps(src + k), b = _mm256_loadu_ps(src + k + 8), c = _mm256_loadu_ps(src + k + 16), d = _mm256_loadu_ps(src + k + 24);
    const __m256i ia = _mm256_cvttps_epi32(a), ib = _mm256_cvttps_epi32(b), ic = _mm256_cvttps_epi32(c), id = _mm256_cvttps_epi32(d);
    // exact iff converting back gives the same float (NaN, fractions and |v| >= 2^31 fail here) ...
    bad = _mm256_or_ps(bad, _mm256_or_ps(_mm256_or_ps(_mm256_cmp_ps(_mm256_cvtepi32_ps(ia), a, _CMP_NEQ_UQ), _mm256_cmp_ps(_mm256_cvtepi32_ps(ib), b, _CMP_NEQ_UQ)),
                                         _mm256_or_ps(_mm256_cmp_ps(_mm256_cvtepi32_ps(ic), c, _CMP_NEQ_UQ), _mm256_cmp_ps(_mm256_cvtepi32_ps(id), d, _CMP_NEQ_UQ))));
    // ... and the integer is in 0..255 (any bit above the low byte set -> out of range, negatives included)
    const __m256i hi = _mm256_or_si256(_mm256_or_si256(ia, ib), _mm256_or_si256(ic, id));
    bad = _mm256_or_ps(bad, _mm256_castsi256_ps(_mm256_andnot_si256(_mm256_set1_epi32(0xFF), hi)));
    const __m256i ab = _mm256_packus_epi32(ia, ib), cd = _mm256_packus_epi32(ic, id);     // per 128-bit lane: a0-3 b0-3 | a4-7 b4-7
    const __m256i q = _mm256_permutevar8x32_epi32(_mm256_packus_epi16(ab, cd), perm);       // lanes: a0-3 b0-3 c0-3 d0-3 | a4-7 b4-7 c4-7 d4-7 -> in order
    _mm256_storeu_si256(reinterpret_cast<__m256i*>(dst + k), q);
    if ((k & 1023) == 992 && !_mm256_testz_si256(_mm256_castps_si256(bad), _mm256_castps_si256(bad))) return false;
  }
  if (!_mm256_testz_si256(_mm256_castps_si256(bad), _mm256_castps_si256(bad))) return false;
  return f32_to_u8_checked_scalar(src + k, dst + k, n - k);
}
#endif

// dst[k] = (uint8_t)src[k] for k < n when EVERY src[k] is an integer in 0..255; returns false (dst unspecified) otherwise.
inline bool f32_to_u8_checked(const float* src, uint8_t* dst, size_t n) {
#if defined(__x86_64__)
  static const bool have_avx2 = __builtin_cpu_supports("avx2");
  if (have_avx2) return f32_to_u8_checked_avx2(src, dst, n);
#endif
  return f32_to_u8_checked_scalar(src, dst, n);
}

}  // namespace b200m
