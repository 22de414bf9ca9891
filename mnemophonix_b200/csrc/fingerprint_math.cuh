// fingerprint_math.cuh — the exact per-pixel and per-Haar-step arithmetic shared by the group-record kernel
// (k15_groups.cu) and the fingerprint kernel (k2_fingerprint.cu).
#pragma once
#include "common.cuh"

namespace mnemo {

// haar.c:35-36: (float)((double)(a +/- b) / M_SQRT2).
// Two exact forms (common.cuh): haar_*64 divides in double by reciprocal and is exact for every float;
// haar_*  stays in FP32 (no conversions, no FP64) and is exact whenever the dividend is 0 or at least 2^-103
// in magnitude.  Where that is guaranteed:
//   pixels are 0 or >= logf(1 + 2^-23) / logf(256) > 2^-26, so every value of the time transform is a sum of
//   non-negative terms (a low: 0 or >= 2^-30, a multiple of 2^-53) or a difference of two lows (0 or >= 2^-53);
//   its outputs are 0 or >= 2^-54, i.e. multiples of 2^-77.  Band level 1 therefore divides 0 or >= 2^-77 and
//   yields multiples of 2^-101; band level 2 divides 0 or >= 2^-101.  From band level 3 on a (wildly
//   adversarial) chain of cancellations could go below 2^-103, so those 3.5 of 15.5 steps keep the double form.
__device__ __forceinline__ float haar_lo(float a, float b) { return div_sqrt2_f32(__fadd_rn(a, b)); }
__device__ __forceinline__ float haar_hi(float a, float b) { return div_sqrt2_f32(__fsub_rn(a, b)); }
// both outputs of one step: lo = (a + b) / sqrt2, hi = (a - b) / sqrt2 (packed FP32 form of haar_lo / haar_hi)
__device__ __forceinline__ void haar_pair(float a, float b, float& lo, float& hi) {
  div_sqrt2_f32x2(__fadd_rn(a, b), __fsub_rn(a, b), lo, hi);
}
__device__ __forceinline__ float haar_lo64(float a, float b) {
  return div_const_f32(__fadd_rn(a, b), kSqrt2, kInvSqrt2);
}
__device__ __forceinline__ float haar_hi64(float a, float b) {
  return div_const_f32(__fsub_rn(a, b), kSqrt2, kInvSqrt2);
}

// spectralimages.c:52-58 for one pixel
template <int LOGF_VARIANT>
__device__ __forceinline__ float scale_pixel(float v, double dmx, double rmx, float log256, float inv_log256,
                                             uint32_t logtab_saddr) {
  float sc = __double2float_rn(div_markstein2(__dmul_rn(255.0, (double)v), dmx, rmx));   // spectralimages.c:53
  if (sc > 255.0f) sc = 255.0f;
  // spectralimages.c:57: an f32 division by logf(256.0f); the dividend is 0 or >= logf(1 + 2^-23) = 1.19e-7,
  // inside the domain where the reciprocal form is exact (common.cuh, mnemo_selftest_div which = 4)
  return div_f32_recip(logf_glibc<LOGF_VARIANT>(__fadd_rn(1.0f, sc), logtab_saddr), log256, inv_log256);
}

}  // namespace mnemo
