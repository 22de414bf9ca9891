// k2_fingerprint.cu — 128x32 window of band energies -> 100-byte MinHash signature.
//
// Replaces the image gather + scale_to_full_spectrum (spectralimages.c:52-77,116-123),
// transform_image (haar.c:23-73), the qsort-based top-200 selection and tri-state bit packing
// (rawfingerprints.c:43-100) and calculate_signature (minhash.c:13-28).  One CTA of 256
// threads owns one spectral image from the (L2-resident) band energies to the signature; the
// image never exists in global memory.
//   1. thread (band = lane, t-segment = warp) loads 16 energies (coalesced 128-byte rows)
//   2. block max; scale = logf(1 + min(255, (float)(255.0*v/max))) / logf(256), the divide in
//      double and logf mirroring glibc's algorithm bit for bit (common.cuh)
//   3. Haar along time: 4 levels in registers, 3 more by warp 0 on the 8 partial lows per band
//   4. Haar along bands: 2 threads per row, 4 levels in registers + 1 shuffle level
//      (every step (float)((double)(a +/- b) / M_SQRT2), as haar.c:35-36)
//   5. exact top-200 by |c| with ties to the lower index (what glibc's stable merge sort gives
//      the reference): 8 rounds of 4-bit radix select on the magnitude bits, counters in
//      registers, then an index-ordered rank among the elements equal to the threshold
//   6. each thread owns 16 consecutive coefficients = exactly one 32-bit word of the
//      fingerprint: bits are assembled without atomics
//   7. MinHash: one warp per permutation, 32 positions per ballot, first set bit wins
#include "common.cuh"

namespace mnemo {

constexpr int kK2Threads = 256;
constexpr int kBStride = 33;

__device__ __forceinline__ float haar_lo(float a, float b) {
  return __double2float_rn(__ddiv_rn((double)__fadd_rn(a, b), 1.41421356237309504880));
}
__device__ __forceinline__ float haar_hi(float a, float b) {
  return __double2float_rn(__ddiv_rn((double)__fsub_rn(a, b), 1.41421356237309504880));
}

template <int LOGF_VARIANT>
__global__ void __launch_bounds__(kK2Threads, 4)
    k2_fingerprint_kernel(const TrackDev* __restrict__ tracks, const uint64_t* __restrict__ img_prefix,
                          uint32_t n_tracks, const Tables* __restrict__ tab, const uint16_t* __restrict__ perms,
                          const float* __restrict__ bins, uint8_t* __restrict__ sig_dense, uint8_t* __restrict__ keep,
                          uint8_t* __restrict__ tap_raw, float* __restrict__ tap_img, float* __restrict__ tap_haar) {
  __shared__ float B[kImgW * kBStride];
  __shared__ float lows[8][kBands];
  __shared__ uint32_t bits[256];
  __shared__ uint32_t hist[8][8];
  __shared__ uint32_t red[8];
  __shared__ uint32_t s_strong, s_any;
  __shared__ uint8_t s_sig[104];

  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const uint64_t img = blockIdx.x;
  const uint32_t t = find_segment<uint64_t>(img_prefix, n_tracks, img);
  const TrackDev tr = tracks[t];
  const uint64_t j = img - img_prefix[t];
  const float* __restrict__ src = bins + (tr.frame_off + j * kImgStep) * kBands;

  if (tid < 64) (&hist[0][0])[tid] = 0;
  if (tid == 0) {
    s_strong = 0;
    s_any = 0;
  }

  // 1. load: element (time = 16*warp + u, band = lane)
  float v[16];
#pragma unroll
  for (int u = 0; u < 16; ++u) v[u] = src[(16 * warp + u) * kBands + lane];

  // 2. max (energies are >= +0, so unsigned bit order is value order)
  uint32_t mb = 0;
#pragma unroll
  for (int u = 0; u < 16; ++u) mb = max(mb, __float_as_uint(v[u]));
  mb = __reduce_max_sync(0xffffffffu, mb);
  if (lane == 0) red[warp] = mb;
  __syncthreads();
  mb = red[0];
#pragma unroll
  for (int w = 1; w < 8; ++w) mb = max(mb, red[w]);
  const float mx = __uint_as_float(mb);
  if (!(mx > 0.0f) || !(mx < __int_as_float(0x7f800000))) {
    // all-zero window: 255*0/0 = NaN everywhere, the reference's sort keeps the first 200
    // (all NaN), sets no bit and flags silence (rawfingerprints.c:64-71,96).  Same for NaN/inf.
    if (tid == 0) keep[img] = 0;
    if (tap_raw) reinterpret_cast<uint32_t*>(tap_raw + img * kRawBytes)[tid] = 0;
    return;
  }
  const double dmx = (double)mx;
  const float log256 = tab->log256;
#pragma unroll
  for (int u = 0; u < 16; ++u) {
    float sc = __double2float_rn(__ddiv_rn(__dmul_rn(255.0, (double)v[u]), dmx));   // spectralimages.c:53
    if (sc > 255.0f) sc = 255.0f;
    v[u] = __fdiv_rn(logf_glibc<LOGF_VARIANT>(__fadd_rn(1.0f, sc)), log256);          // spectralimages.c:57
  }
  if (tap_img) {
#pragma unroll
    for (int u = 0; u < 16; ++u) tap_img[img * kImgElems + (16 * warp + u) * kBands + lane] = v[u];
  }

  // 3. Haar along time (haar.c:56-67): row = band `lane`, this thread holds t in [16*warp, 16*warp+16)
  {
    float l1[8], l2[4], l3[2];
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      l1[i] = haar_lo(v[2 * i], v[2 * i + 1]);
      B[(64 + 8 * warp + i) * kBStride + lane] = haar_hi(v[2 * i], v[2 * i + 1]);
    }
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      l2[i] = haar_lo(l1[2 * i], l1[2 * i + 1]);
      B[(32 + 4 * warp + i) * kBStride + lane] = haar_hi(l1[2 * i], l1[2 * i + 1]);
    }
#pragma unroll
    for (int i = 0; i < 2; ++i) {
      l3[i] = haar_lo(l2[2 * i], l2[2 * i + 1]);
      B[(16 + 2 * warp + i) * kBStride + lane] = haar_hi(l2[2 * i], l2[2 * i + 1]);
    }
    lows[warp][lane] = haar_lo(l3[0], l3[1]);
    B[(8 + warp) * kBStride + lane] = haar_hi(l3[0], l3[1]);
  }
  __syncthreads();
  if (warp == 0) {
    float l4[8], l5[4], l6[2];
#pragma unroll
    for (int i = 0; i < 8; ++i) l4[i] = lows[i][lane];
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      l5[i] = haar_lo(l4[2 * i], l4[2 * i + 1]);
      B[(4 + i) * kBStride + lane] = haar_hi(l4[2 * i], l4[2 * i + 1]);
    }
#pragma unroll
    for (int i = 0; i < 2; ++i) {
      l6[i] = haar_lo(l5[2 * i], l5[2 * i + 1]);
      B[(2 + i) * kBStride + lane] = haar_hi(l5[2 * i], l5[2 * i + 1]);
    }
    B[1 * kBStride + lane] = haar_hi(l6[0], l6[1]);
    B[0 * kBStride + lane] = haar_lo(l6[0], l6[1]);
  }
  __syncthreads();

  // 4. Haar along bands (haar.c:70-72): row = tid/2, this thread holds bands [16*half, 16*half+16)
  const int row = tid >> 1, half = tid & 1;
  float c[16];
  {
    float* r = B + row * kBStride;
    float x[16];
#pragma unroll
    for (int e = 0; e < 16; ++e) x[e] = r[16 * half + e];
    __syncwarp();
    float l1[8], l2[4], l3[2];
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      l1[i] = haar_lo(x[2 * i], x[2 * i + 1]);
      r[16 + 8 * half + i] = haar_hi(x[2 * i], x[2 * i + 1]);
    }
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      l2[i] = haar_lo(l1[2 * i], l1[2 * i + 1]);
      r[8 + 4 * half + i] = haar_hi(l1[2 * i], l1[2 * i + 1]);
    }
#pragma unroll
    for (int i = 0; i < 2; ++i) {
      l3[i] = haar_lo(l2[2 * i], l2[2 * i + 1]);
      r[4 + 2 * half + i] = haar_hi(l2[2 * i], l2[2 * i + 1]);
    }
    const float l4 = haar_lo(l3[0], l3[1]);
    r[2 + half] = haar_hi(l3[0], l3[1]);
    const float other = __shfl_xor_sync(0xffffffffu, l4, 1);
    const float a = half ? other : l4, b = half ? l4 : other;
    r[half] = half ? haar_hi(a, b) : haar_lo(a, b);
    __syncwarp();
#pragma unroll
    for (int e = 0; e < 16; ++e) c[e] = r[16 * half + e];   // final coefficients 16*tid .. 16*tid+15
  }
  if (tap_haar) {
#pragma unroll
    for (int e = 0; e < 16; ++e) tap_haar[img * kImgElems + 16 * tid + e] = c[e];
  }

  // 5. exact 200th largest magnitude: radix select, 4 bits per round
  uint32_t key[16];
#pragma unroll
  for (int e = 0; e < 16; ++e) key[e] = __float_as_uint(c[e]) & 0x7fffffffu;
  uint32_t prefix = 0, decided = 0, want = kTopK, n_eq = 0;
#pragma unroll 1
  for (int it = 0; it < 8; ++it) {
    const int shift = 28 - 4 * it;
    uint64_t c0 = 0, c1 = 0;   // 16 byte-wide counters (<= 16 each)
#pragma unroll
    for (int e = 0; e < 16; ++e) {
      const bool in = (key[e] & decided) == prefix;
      const uint32_t d = (key[e] >> shift) & 15u;
      const uint64_t inc = in ? (1ull << (8 * (d & 7u))) : 0ull;
      c0 += (d < 8u) ? inc : 0ull;
      c1 += (d < 8u) ? 0ull : inc;
    }
    // widen to 16-bit fields, two digits per word, and reduce across the warp
#pragma unroll
    for (int w = 0; w < 8; ++w) {
      const uint64_t s = w < 4 ? c0 : c1;
      const int sh = 16 * (w & 3);
      uint32_t pair = (uint32_t)((s >> sh) & 0xffu) | ((uint32_t)((s >> (sh + 8)) & 0xffu) << 16);
      pair = __reduce_add_sync(0xffffffffu, pair);
      if (lane == 0) atomicAdd(&hist[it][w], pair);
    }
    __syncthreads();
    uint32_t above = 0;
    int dsel = 0;
    uint32_t cnt_sel = 0;
    bool found = false;
#pragma unroll
    for (int d = 15; d >= 0; --d) {
      const uint32_t cnt = (hist[it][d >> 1] >> (16 * (d & 1))) & 0xffffu;
      if (!found && above + cnt >= want) {
        found = true;
        dsel = d;
        cnt_sel = cnt;
      }
      if (!found) above += cnt;
    }
    prefix |= (uint32_t)dsel << shift;
    decided |= 15u << shift;
    want -= above;
    n_eq = cnt_sel;
  }
  const uint32_t thr = prefix;   // the 200th largest magnitude; `want` of the n_eq equal elements are taken

  // rank among equal elements in index order (stable sort => lower index first)
  uint32_t eq_before = 0;
  if (n_eq > want) {
    uint32_t mine = 0;
#pragma unroll
    for (int e = 0; e < 16; ++e) mine += key[e] == thr;
    uint32_t incl = mine;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const uint32_t y = __shfl_up_sync(0xffffffffu, incl, o);
      if (lane >= o) incl += y;
    }
    if (lane == 31) red[warp] = incl;
    __syncthreads();
    uint32_t base = 0;
#pragma unroll
    for (int w = 0; w < 8; ++w) base += w < warp ? red[w] : 0u;
    eq_before = base + incl - mine;
  }

  // 6. tri-state bits (rawfingerprints.c:61-74); thread tid owns word tid of the 8192-bit array
  uint32_t word = 0, strong = 0;
#pragma unroll
  for (int e = 0; e < 16; ++e) {
    bool sel = key[e] > thr;
    if (key[e] == thr) {
      sel = eq_before < want;
      ++eq_before;
    }
    if (sel) {
      const double cd = (double)c[e];
      if (cd > 0.001) word |= 1u << (2 * e);
      else if (cd < -0.001) word |= 1u << (2 * e + 1);
      strong += (double)fabsf(c[e]) > 1.0;
    }
  }
  bits[tid] = word;
  strong = __reduce_add_sync(0xffffffffu, strong);
  if (lane == 0 && strong) atomicAdd(&s_strong, strong);
  __syncthreads();
  const bool silent = s_strong < 10u;   // rawfingerprints.c:96
  if (tap_raw) reinterpret_cast<uint32_t*>(tap_raw + img * kRawBytes)[tid] = word;
  if (silent) {
    if (tid == 0) keep[img] = 0;
    return;
  }

  // 7. MinHash (minhash.c:13-28)
  for (int p = warp; p < kSigLen; p += 8) {
    const uint16_t* __restrict__ prow = perms + p * kPermLen;
    uint32_t res = kPermLen;
#pragma unroll 1
    for (int cidx = 0; cidx < 8; ++cidx) {
      const int jj = 32 * cidx + lane;
      bool hit = false;
      if (jj < kPermLen) {
        const uint32_t pos = __ldg(prow + jj);
        hit = (bits[pos >> 5] >> (pos & 31u)) & 1u;
      }
      const uint32_t bal = __ballot_sync(0xffffffffu, hit);
      if (bal) {
        res = 32 * cidx + __ffs(bal) - 1;
        break;
      }
    }
    if (lane == 0) {
      s_sig[p] = (uint8_t)res;
      if (res != kPermLen) s_any = 1;
    }
  }
  __syncthreads();
  if (tid < 25) reinterpret_cast<uint32_t*>(sig_dense + img * kSigLen)[tid] = reinterpret_cast<uint32_t*>(s_sig)[tid];
  if (tid == 0) keep[img] = s_any ? 1 : 0;
}

void launch_k2_fingerprint(const TrackDev* tracks, const uint64_t* img_prefix, uint32_t n_tracks, uint64_t n_images,
                           const Tables* tab, const uint16_t* perms, const float* bins, uint8_t* sig_dense,
                           uint8_t* keep, uint8_t* tap_raw, float* tap_img, float* tap_haar, int logf_variant,
                           cudaStream_t st) {
  if (n_images == 0) return;
  if (logf_variant)
    k2_fingerprint_kernel<1><<<(unsigned)n_images, kK2Threads, 0, st>>>(tracks, img_prefix, n_tracks, tab, perms, bins,
                                                                        sig_dense, keep, tap_raw, tap_img, tap_haar);
  else
    k2_fingerprint_kernel<0><<<(unsigned)n_images, kK2Threads, 0, st>>>(tracks, img_prefix, n_tracks, tab, perms, bins,
                                                                        sig_dense, keep, tap_raw, tap_img, tap_haar);
}

// ---- device self-test of the logf mirror: compare against host values for a bit range ----
__global__ void logf_range_kernel(uint32_t lo_bits, uint32_t n, int variant, float* __restrict__ out) {
  const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const float x = __uint_as_float(lo_bits + i);
  out[i] = variant ? logf_glibc<1>(x) : logf_glibc<0>(x);
}

void launch_logf_range(uint32_t lo_bits, uint32_t n, int variant, float* out, cudaStream_t st) {
  if (n) logf_range_kernel<<<(n + 255) / 256, 256, 0, st>>>(lo_bits, n, variant, out);
}

}  // namespace mnemo
