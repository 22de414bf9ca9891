// common.cuh — shared device/host declarations for libmnemo_cuda.so (sm_100a only).
//
// Arithmetic contract: every float stage reproduces the reference's C arithmetic exactly
// (SURVEY.md App. A): separate IEEE mul/add roundings (no FMA contraction; the reference is
// built for baseline x86-64), double-precision islands where C promotes to double, and
// denormals kept.  Build flags: -fmad=false -ftz=false -prec-div=true -prec-sqrt=true.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

// Checked build (-DMNEMO_CHECKED, tools/checked_build.sh): bounds assertions on the data-dependent indices of the search
// kernels; a violation prints its condition and traps, which fails whatever GPU test was running.  compute-sanitizer is
// closed on the development pool, so this is the memory-safety evidence beside bit-exact parity.  Without the flag
// the macros vanish (identical SASS).
#ifdef MNEMO_CHECKED
#include <cstdio>
#define MN_CHECK(cond)                                                                                              \
  do {                                                                                                              \
    if (!(cond)) {                                                                                                  \
      printf("MN_CHECK failed: %s at %s:%d (block %u thread %u)\n", #cond, __FILE__, __LINE__, blockIdx.x, threadIdx.x); \
      __trap();                                                                                                     \
    }                                                                                                               \
  } while (0)
#else
#define MN_CHECK(cond) \
  do {                 \
  } while (0)
#endif

namespace mnemo {

constexpr int kFrame = 2048;        // spectralimages.h:12
constexpr int kHop = 64;            // spectralimages.h:18
constexpr int kBands = 32;          // logbins.h:5
constexpr int kImgW = 128;          // spectralimages.h:24
constexpr int kImgStep = 8;         // spectralimages.h:30
constexpr int kImgElems = kImgW * kBands;
constexpr int kTopK = 200;          // rawfingerprints.h:13
constexpr int kRawBytes = 1024;     // rawfingerprints.h:16
constexpr int kSigLen = 100;        // minhash.h:7
constexpr int kPermLen = 255;       // permutations.h:10
constexpr int kTaps = 31;           // resample.c:7
constexpr int kDecim = 8;
constexpr int kTables = 25;         // lsh.h:9

constexpr int kSumChunk = 4096;     // samples per speculative sum-of-squares chunk
constexpr int kSumCands = 3;        // candidate binades simulated per chunk
constexpr uint32_t kSumInvalid = 0xFFFFFFFFu;

// Per-track descriptor (device copy lives in the batch).
struct TrackDev {
  const int16_t* pcm;     // device pointer, interleaved int16 (nullptr for pre-normalised input)
  uint64_t n_pcm;         // sample frames
  uint32_t channels;
  uint32_t n5512;         // n_pcm / 8
  uint32_t n_frames;      // 0 when the track is too small to fingerprint
  uint32_t n_images;
  uint32_t prenormalized; // samples are already normalised: skip rms division (from_samples path)
  uint32_t n_chunks;      // ceil(n5512 / kSumChunk)
  uint64_t s_off;         // float offset of this track's 5512 Hz samples
  uint64_t frame_off;     // frame offset into bins
  uint64_t img_off;       // image offset into per-image arrays
  uint64_t chunk_off;     // chunk offset into the sum-of-squares scratch
  uint64_t group_off;     // offset of this track's 8-frame groups (n_frames / 8 of them) into the group arrays
};

// Host-computed tables (same libm calls as the reference), uploaded once per context.
struct Tables {
  float taps[kTaps];            // resample.c:27-35
  float hann[kFrame];           // hannwindow.c:8-12
  float2 tw64[32];              // twiddles W64^m, m<32 = table[32m]        (fft.c:49-51)
  float2 tw_hi[2][31][32];      // stage 7..11 twiddles per (half, slot, lane)
  uint16_t edges[kBands + 1];   // logbins.c:44-55
  float log256;                 // logf(256.0f)                              (spectralimages.c:57)
  float inv_log256;             // RN(1 / log256) in f32, for the exact reciprocal division
  int logf_variant;             // which glibc logf build the host runs (0 baseline, 1 fma)
};

struct SumChunkRec {   // one per chunk, 32 bytes
  uint32_t e_hi;       // biased exponent of the highest candidate binade (0 = no candidates)
  uint32_t delta[kSumCands][2];  // mantissa increment for (candidate, parity of incoming mantissa)
  uint32_t pad;
};

// ---- launchers (defined in the k*.cu files); all asynchronous on `st` ----
void launch_k0_resample(const TrackDev* tracks, const uint32_t* tile_prefix, uint32_t n_tracks, uint32_t n_tiles,
                        const Tables* tab, float* s5512, cudaStream_t st);
void launch_sumsq(const TrackDev* tracks, uint32_t n_tracks, uint64_t n_chunks_total, const uint32_t* chunk_track,
                  const float* s5512, double* chunk_true, double* chunk_prefix, double* chunk_prefix_end,
                  SumChunkRec* recs, SumChunkRec* grecs, uint32_t* slice_maps, float* sumsq, float* rms, cudaStream_t st,
                  int* n_launches);
void launch_normalize(const TrackDev* tracks, const uint32_t* chunk_track, uint64_t n_chunks, const float* s5512,
                      const float* rms, float* snorm, cudaStream_t st);
void launch_k1_spectral(const TrackDev* tracks, const uint32_t* tile_prefix, uint32_t n_tracks, uint32_t n_tiles,
                        const Tables* tab, const float* snorm, float* bins, int sm_count, cudaStream_t st);
void launch_group_records(const TrackDev* tracks, const uint32_t* group_prefix, uint32_t n_tracks, uint32_t n_groups,
                          const Tables* tab, const float* bins, uint32_t* gmax, float* imax, uint8_t* slot, float* rec,
                          float* tap_px, int logf_variant, cudaStream_t st, int* n_launches);
void launch_k2_fingerprint(const TrackDev* tracks, const uint64_t* img_prefix, uint32_t n_tracks, uint64_t n_images,
                           const uint16_t* inv_perms, const float* imax, const uint8_t* slot, const float* rec,
                           const float* tap_px, uint8_t* sig_dense, uint8_t* keep, uint8_t* tap_raw, float* tap_img,
                           float* tap_haar, cudaStream_t st);
void launch_compact(const uint8_t* sig_dense, const uint8_t* keep, uint64_t n_images, const uint64_t* img_prefix,
                    uint32_t n_tracks, uint32_t* block_sums, uint32_t* pos, uint8_t* sig_out, uint32_t* n_sigs,
                    cudaStream_t st, int* n_launches);
void k1_upload_constants(const Tables& t);
// tolerance mode of K1 (k1_spectral.cu, "fast K1"): twiddle tables from double precision, kernel launch
size_t k1_fast_table_float2();
void k1_fast_tables(float2* tw, float2 w32[16]);
void k1_fast_upload_constants(const float2 w32[16]);
void launch_k1_fast(const TrackDev* tracks, const uint32_t* tile_prefix, uint32_t n_tracks, uint32_t n_tiles, const Tables* tab,
                    const float2* tw_fast, const float* snorm, float* bins, int sm_count, cudaStream_t st);
void k0_upload_constants(const Tables& t);
void launch_div_test(int which, uint32_t lo_bits, uint64_t n, unsigned long long* bad, float c, float rc,
                     cudaStream_t st);
void launch_logf_range(uint32_t lo_bits, uint32_t n, int variant, float* out, cudaStream_t st);
int k1_tile_frames();
int k1_max_band_width();
int k0_tile_outputs();

// ---- device helpers ----
#ifdef __CUDACC__
// glibc 2.39 logf (sysdeps/ieee754/flt-32/e_logf.c, ARM optimized-routines algorithm): 16-entry
// table, degree-3 polynomial in double.  Constants are the published table (identical to the
// .rodata of this image's libm.so.6).  Valid for normal positive finite x (here x in [1, 256]).
static __device__ __constant__ double c_logf_tab[16][2] = {
    {0x1.661ec79f8f3bep+0, -0x1.57bf7808caadep-2}, {0x1.571ed4aaf883dp+0, -0x1.2bef0a7c06ddbp-2},
    {0x1.49539f0f010bp+0, -0x1.01eae7f513a67p-2},  {0x1.3c995b0b80385p+0, -0x1.b31d8a68224e9p-3},
    {0x1.30d190c8864a5p+0, -0x1.6574f0ac07758p-3}, {0x1.25e227b0b8eap+0, -0x1.1aa2bc79c81p-3},
    {0x1.1bb4a4a1a343fp+0, -0x1.a4e76ce8c0e5ep-4}, {0x1.12358f08ae5bap+0, -0x1.1973c5a611cccp-4},
    {0x1.0953f419900a7p+0, -0x1.252f438e10c1ep-5}, {0x1p+0, 0x0p+0},
    {0x1.e608cfd9a47acp-1, 0x1.aa5aa5df25984p-5},  {0x1.ca4b31f026aap-1, 0x1.c5e53aa362eb4p-4},
    {0x1.b2036576afce6p-1, 0x1.526e57720db08p-3},  {0x1.9c2d163a1aa2dp-1, 0x1.bc2860d22477p-3},
    {0x1.886e6037841edp-1, 0x1.1058bc8a07ee1p-2},  {0x1.767dcf5534862p-1, 0x1.4043057b6ee09p-2}};

// Polynomial coefficients and ln2 live in the constant bank so that the FP64 instructions take them as
// operands (as immediates they cost two UMOVs each, re-materialised for every pixel).
static __device__ __constant__ double c_logf_poly[4] = {-0x1.00ea348b88334p-2, 0x1.5575b0be00b6ap-2,
                                                         -0x1.ffffef20a4123p-2, 0x1.62e42fefa39efp-1};   // A0 A1 A2 Ln2

// tab_saddr: shared-window address of a copy of the table split into two 16-entry double arrays,
// invc[16] followed by logc[16] (256 bytes, 128-byte aligned).  The index is data dependent, and 16 doubles
// span exactly the 32 banks, so a 64-bit read of either array is conflict-free for any index pattern;
// the two reads share one address computation.  tab_saddr == 0 reads the constant table instead.
// glibc returns +0 for x == 1 through an early exit; the general path gives the same +0 (r = 0, k = 0,
// logc[9] = 0), so there is no branch here (the exhaustive self-test covers x = 1).
template <int VARIANT>
__device__ __forceinline__ float logf_glibc(float x, uint32_t tab_saddr = 0) {
  const double a0 = c_logf_poly[0], a1 = c_logf_poly[1], a2 = c_logf_poly[2], ln2 = c_logf_poly[3];
  const uint32_t ix = __float_as_uint(x);
  const uint32_t tmp = ix - 0x3f330000u;
  const int k = (int)tmp >> 23;
  const uint32_t iz = ix - (tmp & 0xff800000u);
  const double z = (double)__uint_as_float(iz);
  double invc, logc;
  if (tab_saddr) {
    const uint32_t a = tab_saddr + ((tmp >> 16) & 0x78u);   // 8 * ((tmp >> 19) & 15)
    asm("ld.shared.f64 %0, [%2];\n\tld.shared.f64 %1, [%2+128];" : "=d"(invc), "=d"(logc) : "r"(a));
  } else {
    const int i = (tmp >> 19) & 15;
    invc = c_logf_tab[i][0];
    logc = c_logf_tab[i][1];
  }
  if (VARIANT) {  // instruction order of the -mfma ifunc variant
    double y0 = __fma_rn((double)k, ln2, logc);
    double r = __fma_rn(z, invc, -1.0);
    double y = __fma_rn(r, a1, a2);
    double r2 = __dmul_rn(r, r);
    double hi = __dadd_rn(r, y0);
    y = __fma_rn(r2, a0, y);
    return __double2float_rn(__fma_rn(r2, y, hi));
  } else {
    double r = __dadd_rn(__dmul_rn(z, invc), -1.0);
    double y0 = __dadd_rn(logc, __dmul_rn((double)k, ln2));
    double r2 = __dmul_rn(r, r);
    double y = __dadd_rn(__dmul_rn(a1, r), a2);
    y = __dadd_rn(__dmul_rn(a0, r2), y);
    return __double2float_rn(__dadd_rn(__dmul_rn(y, r2), __dadd_rn(y0, r)));
  }
}

// ---- exact division by reciprocal (Markstein) ----
// q = RN(a / b) in double from y = RN(1 / b): q0 = RN(a*y), r0 = a - b*q0 (exact, one FMA),
// q1 = RN(q0 + r0*y).  One more residual step gives the correctly rounded quotient whenever q1 is
// a faithful rounding (Markstein's theorem), which it always is here.  For the two constant
// divisors of the path (sqrt(2) in haar.c:35-36, 32767 in wav.c:373) the ONE-step form is verified
// against IEEE division for every float dividend by mnemo_selftest_div (tests/test_gpu_index.py);
// the per-image divisor of spectralimages.c:53 uses the two-step form.
__device__ __forceinline__ double div_markstein1(double a, double b, double y) {
  const double q0 = __dmul_rn(a, y);
  const double r0 = __fma_rn(-b, q0, a);
  return __fma_rn(r0, y, q0);
}
__device__ __forceinline__ double div_markstein2(double a, double b, double y) {
  const double q1 = div_markstein1(a, b, y);
  const double r1 = __fma_rn(-b, q1, a);
  return __fma_rn(r1, y, q1);
}
// (float)((double)x / b) for the constant divisors b > 0, one-step form.  mnemo_selftest_div checks all
// 2^32 float dividends against __ddiv_rn: the only difference is x = -0, which comes back as +0.  A -0
// dividend cannot occur on this path (the dividends are (float)int / channels in K0, and sums/differences
// of values that are never -0 in the Haar steps), and the sign of a zero is not observable downstream
// (|c| keys, c > 0.001 tests, s*s); the self-test therefore skips that single input.
__device__ __forceinline__ float div_const_f32(float x, double b, double y) {
  return __double2float_rn(div_markstein1((double)x, b, y));
}
constexpr double kSqrt2 = 1.41421356237309504880;        // M_SQRT2
constexpr double kInvSqrt2 = 0.70710678118654746172;     // RN(1 / M_SQRT2) as a double

// ---- the same constant divisions without leaving FP32 ----
// (float)((double)x / d) costs two conversions on the 16-lane XU pipe plus three FP64 instructions; the XU pipe
// was the busiest pipe of K0 and K2.  With d split into two floats h + l and y = RN32(1 / d):
//   q0 = RN(x * y);  r = x - q0 * (h + l) by two FMAs (the first one exact);  q1 = RN(q0 + r * y)
// is the correctly rounded quotient, and rounding through double first never changes it.  Not taken on trust:
// mnemo_selftest_div compares each form against the double-precision division for EVERY float dividend
// (tools/div_f32_probe.cu is the search that found the domains):
//   x / 32767.0       exact for all 2^32 floats (32767 is a float, l = 0)                      -> K0
//   x / M_SQRT2       exact for x == 0 or |x| >= 2^-103 (below that the residual underflows)    -> K2, see haar_*()
//   x / logf(256.0f)  (an f32 division in the reference) exact for x == 0 or |x| >= 2^-106      -> K2 pixel scale
// As with div_const_f32, a -0 dividend comes back as +0 and cannot occur.
constexpr float kInv32767f = 0x1.0002p-15f;              // RN32(1 / 32767)
constexpr float kSqrt2Hi = 0x1.6a09e6p+0f;               // RN32(M_SQRT2)
constexpr float kSqrt2Lo = 0x1.9fcef4p-26f;              // RN32(M_SQRT2 - kSqrt2Hi)
constexpr float kInvSqrt2f = 0x1.6a09e6p-1f;             // RN32(1 / M_SQRT2)
constexpr float kDivSqrt2MinAbs = 0x1p-103f;
constexpr float kDivLog256MinAbs = 0x1p-106f;

// x * SCALE / 32767.0 with SCALE a power of two folded into the constants (SCALE = 0.5: the stereo down-mix
// (float)sum / 2.0f / 32767.0 in one go; scaling by two is exact, so every rounding happens at the same place)
template <int HALVE>
__device__ __forceinline__ float div32767_f32(float x) {
  const float y = HALVE ? 0.5f * kInv32767f : kInv32767f;
  const float h = HALVE ? 2.0f * 32767.0f : 32767.0f;
  const float q0 = __fmul_rn(x, y);
  const float r = __fmaf_rn(-h, q0, x);
  return __fmaf_rn(r, y, q0);
}
__device__ __forceinline__ float div_sqrt2_f32(float x) {
  const float q0 = __fmul_rn(x, kInvSqrt2f);
  const float r = __fmaf_rn(-kSqrt2Lo, q0, __fmaf_rn(-kSqrt2Hi, q0, x));
  return __fmaf_rn(r, kInvSqrt2f, q0);
}
// Two such divisions at once with Blackwell's packed FP32 instructions (FMUL2 / FFMA2: two independent
// IEEE-rounded operations per issue slot), for the (a + b, a - b) pair of a Haar step: 4 instructions
// instead of 8, bit-identical to div_sqrt2_f32 on each half.
__device__ __forceinline__ uint64_t pack_f32x2(float x, float y) {
  uint64_t r;
  asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(x), "f"(y));
  return r;
}
__device__ __forceinline__ void div_sqrt2_f32x2(float x0, float x1, float& q0_out, float& q1_out) {
  const uint64_t y = pack_f32x2(kInvSqrt2f, kInvSqrt2f), nh = pack_f32x2(-kSqrt2Hi, -kSqrt2Hi),
                 nl = pack_f32x2(-kSqrt2Lo, -kSqrt2Lo);
  const uint64_t x = pack_f32x2(x0, x1);
  uint64_t q0, r, q1;
  asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(q0) : "l"(x), "l"(y));
  asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(r) : "l"(nh), "l"(q0), "l"(x));
  asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(r) : "l"(nl), "l"(q0), "l"(r));
  asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(q1) : "l"(r), "l"(y), "l"(q0));
  asm("mov.b64 {%0, %1}, %2;" : "=f"(q0_out), "=f"(q1_out) : "l"(q1));
}
// x / c for an f32 constant c with y = RN32(1 / c)
__device__ __forceinline__ float div_f32_recip(float x, float c, float y) {
  const float q0 = __fmul_rn(x, y);
  const float r = __fmaf_rn(-c, q0, x);
  return __fmaf_rn(r, y, q0);
}

// index of the last prefix[] entry <= v (prefix has n+1 ascending entries, prefix[0] == 0)
template <typename T>
__device__ __forceinline__ uint32_t find_segment(const T* __restrict__ prefix, uint32_t n, T v) {
  uint32_t lo = 0, hi = n;
  while (hi - lo > 1) {
    uint32_t mid = (lo + hi) >> 1;
    if (prefix[mid] <= v) lo = mid;
    else hi = mid;
  }
  return lo;
}
#endif

}  // namespace mnemo
