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
//      the reference): bisection of the key range with block-wide counts over the register-resident
//      coefficients until <= 32 keys remain, those ranked by one warp, then an index-ordered rank
//      among the elements equal to the threshold
//   6. each thread owns 16 consecutive coefficients = exactly one 32-bit word of the
//      fingerprint: bits are assembled without atomics
//   7. MinHash inside out: each set bit contributes its position in every permutation (inverse table,
//      100 x u16 per bit), signature = element-wise minimum (VIMNMX3.U16x2)
#include "fingerprint_math.cuh"

namespace mnemo {

constexpr int kK2Threads = 256;
constexpr int kBStride = 33;

// 16-bit two-lane minimum of three (one VIMNMX3.U16x2)
__device__ __forceinline__ uint32_t min3_u16x2(uint32_t a, uint32_t b, uint32_t c) { return __vimin3_u16x2(a, b, c); }

// TAPS: also write the scaled image, the Haar coefficients and the raw fingerprint to global memory (stage
// parity tests only; a separate instantiation so that the production kernel carries none of it)
template <bool TAPS>
__global__ void __launch_bounds__(kK2Threads, 5)
    k2_fingerprint_kernel(const TrackDev* __restrict__ tracks, const uint64_t* __restrict__ img_prefix,
                          uint32_t n_tracks, const uint16_t* __restrict__ inv, const float* __restrict__ imax,
                          const uint8_t* __restrict__ slot, const float* __restrict__ rec,
                          const float* __restrict__ tap_px, uint8_t* __restrict__ sig_dense,
                          uint8_t* __restrict__ keep, uint8_t* __restrict__ tap_raw, float* __restrict__ tap_img,
                          float* __restrict__ tap_haar) {
  __shared__ float B[kImgW * kBStride];
  __shared__ float lows[8][kBands];
  __shared__ __align__(16) uint32_t s_list[208];   // inverse-table rows of the set bits (<= 200), padded to 4
  __shared__ uint2 s_part[8][32];           // per-warp partial MinHash minima
  __shared__ uint32_t s_small[32];
  __shared__ uint32_t s_sel[3], s_count[2];
  __shared__ uint32_t red[8];
  __shared__ __align__(16) uint32_t s_cnt[4][8];   // per-warp maxima exchange
  __shared__ uint32_t s_tot[3][2];                 // rotating block totals of the selection steps
  __shared__ uint32_t s_strong;

  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const uint64_t img = blockIdx.x;
  const uint32_t t = find_segment<uint64_t>(img_prefix, n_tracks, img);
  const TrackDev tr = tracks[t];
  const uint64_t j = img - img_prefix[t];
  if (tid == 0) s_strong = 0;

  // 1.-3. The image's maximum, its scaled pixels and the first three Haar levels along time come from the
  // group records (k15_groups.cu): image j = frame groups j .. j+15, warp w takes groups 2w and 2w+1 (frames
  // 16w .. 16w+15), lane = band.  A record holds, per band, {hi1[0..3]; hi2[0..1], hi3, lo3} of its 8 frames.
  const float mx = imax[img];
  if (!(mx > 0.0f) || !(mx < __int_as_float(0x7f800000))) {
    // all-zero window: 255*0/0 = NaN everywhere, the reference's sort keeps the first 200
    // (all NaN), sets no bit and flags silence (rawfingerprints.c:64-71,96).  Same for NaN/inf.
    if (tid == 0) keep[img] = 0;
    if (TAPS && tap_raw) reinterpret_cast<uint32_t*>(tap_raw + img * kRawBytes)[tid] = 0;
    return;
  }
  {
    float l3[2];
#pragma unroll
    for (int kk = 0; kk < 2; ++kk) {
      const uint32_t k = 2 * warp + kk;
      const uint32_t d = slot[img * 16 + k];
      const uint64_t base = ((tr.group_off + j + k) * 16 + d) * 256;
      const float4* __restrict__ r4 = reinterpret_cast<const float4*>(rec + base + 8 * lane);
      const float4 h1 = __ldg(r4), o = __ldg(r4 + 1);
      B[(64 + 8 * warp + 4 * kk + 0) * kBStride + lane] = h1.x;
      B[(64 + 8 * warp + 4 * kk + 1) * kBStride + lane] = h1.y;
      B[(64 + 8 * warp + 4 * kk + 2) * kBStride + lane] = h1.z;
      B[(64 + 8 * warp + 4 * kk + 3) * kBStride + lane] = h1.w;
      B[(32 + 4 * warp + 2 * kk + 0) * kBStride + lane] = o.x;
      B[(32 + 4 * warp + 2 * kk + 1) * kBStride + lane] = o.y;
      B[(16 + 2 * warp + kk) * kBStride + lane] = o.z;
      l3[kk] = o.w;
      if (TAPS && tap_img && tap_px) {
#pragma unroll
        for (int f = 0; f < 8; ++f)
          tap_img[img * kImgElems + (16 * warp + 8 * kk + f) * kBands + lane] = tap_px[base + f * kBands + lane];
      }
    }
    float l4, h;
    haar_pair(l3[0], l3[1], l4, h);   // level 4: the first step that crosses a group boundary
    lows[warp][lane] = l4;
    B[(8 + warp) * kBStride + lane] = h;
  }
  __syncthreads();
  if (warp == 0) {
    float l4[8], l5[4], l6[2];
#pragma unroll
    for (int i = 0; i < 8; ++i) l4[i] = lows[i][lane];
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      float h;
      haar_pair(l4[2 * i], l4[2 * i + 1], l5[i], h);
      B[(4 + i) * kBStride + lane] = h;
    }
#pragma unroll
    for (int i = 0; i < 2; ++i) {
      float h;
      haar_pair(l5[2 * i], l5[2 * i + 1], l6[i], h);
      B[(2 + i) * kBStride + lane] = h;
    }
    float l7, h7;
    haar_pair(l6[0], l6[1], l7, h7);
    B[1 * kBStride + lane] = h7;
    B[0 * kBStride + lane] = l7;
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
      float h;
      haar_pair(x[2 * i], x[2 * i + 1], l1[i], h);
      r[16 + 8 * half + i] = h;
    }
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      float h;
      haar_pair(l1[2 * i], l1[2 * i + 1], l2[i], h);
      r[8 + 4 * half + i] = h;
    }
#pragma unroll
    for (int i = 0; i < 2; ++i) {   // band levels 3..5: double form (see haar_lo64)
      l3[i] = haar_lo64(l2[2 * i], l2[2 * i + 1]);
      r[4 + 2 * half + i] = haar_hi64(l2[2 * i], l2[2 * i + 1]);
    }
    const float l4 = haar_lo64(l3[0], l3[1]);
    r[2 + half] = haar_hi64(l3[0], l3[1]);
    const float other = __shfl_xor_sync(0xffffffffu, l4, 1);
    const float a = half ? other : l4, b = half ? l4 : other;
    r[half] = half ? haar_hi64(a, b) : haar_lo64(a, b);
    __syncwarp();
#pragma unroll
    for (int e = 0; e < 16; ++e) c[e] = r[16 * half + e];   // final coefficients 16*tid .. 16*tid+15
  }
  if (TAPS && tap_haar) {
#pragma unroll
    for (int e = 0; e < 16; ++e) tap_haar[img * kImgElems + 16 * tid + e] = c[e];
  }

  // 5. exact 200th largest magnitude, by bisection of the key range with the coefficients staying in registers.
  //  (a) the range is [0, largest key + 1); the first trial value is the smallest of the 8 per-warp maxima
  //      (warps own different Haar scales, so that is a cheap guess just below the answer: typically
  //      500-1000 keys lie above it);
  //  (b) each step counts the keys >= three trial values (48 compares per thread, warp reductions, shared
  //      atomics, one barrier) and keeps the quarter of [lo, hi) with count(>= lo) >= 200 > count(>= hi); a warp
  //      whose largest key is below the lowest trial value skips the compares.  About three steps leave 32 or
  //      fewer keys inside [lo, hi);
  //  (c) those are appended to a 32-entry list and ranked by one warp.
  // (Earlier versions: a bitonic sort of the thread maxima for a guaranteed lower bound, a gather of the ~28 % of
  // keys above it into a shared list and 8-bit radix passes with match.any-aggregated histogram atomics:
  // 6 k warp instructions per image against 2.5 k here.)
  uint32_t key[16];
#pragma unroll
  for (int e = 0; e < 16; ++e) key[e] = __float_as_uint(c[e]) & 0x7fffffffu;
  uint32_t lo = 0, hi, wmax, first;
  {
    uint32_t v = key[0];
#pragma unroll
    for (int e = 1; e < 16; ++e) v = max(v, key[e]);
    wmax = __reduce_max_sync(0xffffffffu, v);
    if (lane == 0) s_cnt[1][warp] = wmax;
    if (tid == 0) {
      s_count[0] = 0;
      s_tot[0][0] = s_tot[0][1] = 0;
    }
    __syncthreads();
    const uint4 m0 = *reinterpret_cast<const uint4*>(&s_cnt[1][0]), m1 = *reinterpret_cast<const uint4*>(&s_cnt[1][4]);
    first = min(min(min(m0.x, m0.y), min(m0.z, m0.w)), min(min(m1.x, m1.y), min(m1.z, m1.w)));
    hi = max(max(max(m0.x, m0.y), max(m0.z, m0.w)), max(max(m1.x, m1.y), max(m1.z, m1.w))) + 1u;
  }
  uint32_t c_lo = kImgElems, c_hi = 0;   // count(>= 0) = 4096
  // Four-way steps: three trial values per step (counts c1 >= c2 >= c3), one barrier per step, ~3 steps per
  // image instead of ~6.3 two-way steps.  Warp totals go through shared atomics into one of three rotating slots
  // (the slot of the next step is cleared before this step's barrier); c1 and c2 travel packed in one word
  // (a warp holds 512 keys, the block 4096: 16 bits each are enough).
  uint32_t slot_cur = 0;
  for (int step = 0; hi - lo > 1u && c_lo - c_hi > 32u; ++step) {
    uint32_t p1, p2, p3;
    if (step == 0 && first > lo && first < hi) {
      const uint32_t r = hi - first;
      p1 = first;
      p2 = first + r / 3u;
      p3 = first + (2u * r) / 3u;
    } else {
      const uint32_t q = (hi - lo) >> 2;
      p1 = q ? lo + q : lo + ((hi - lo) >> 1);
      p2 = q ? lo + 2u * q : p1;
      p3 = q ? lo + 3u * q : p1;
    }
    if (wmax >= p1) {                    // warp-uniform; otherwise this warp contributes nothing to any count
      uint32_t n1 = 0, n2 = 0, n3 = 0;
#pragma unroll
      for (int e = 0; e < 16; ++e) {
        n1 += key[e] >= p1;
        n2 += key[e] >= p2;
        n3 += key[e] >= p3;
      }
      uint32_t w0 = n1 | (n2 << 16), w1 = n3;
      asm("redux.sync.add.u32 %0, %0, 0xffffffff;" : "+r"(w0));
      asm("redux.sync.add.u32 %0, %0, 0xffffffff;" : "+r"(w1));
      if (lane == 0) {
        atomicAdd(&s_tot[slot_cur][0], w0);
        atomicAdd(&s_tot[slot_cur][1], w1);
      }
    }
    const uint32_t slot_next = slot_cur == 2 ? 0 : slot_cur + 1;
    if (tid == 0) s_tot[slot_next][0] = s_tot[slot_next][1] = 0;
    __syncthreads();
    const uint32_t t0 = s_tot[slot_cur][0], c3 = s_tot[slot_cur][1];
    const uint32_t c1 = t0 & 0xffffu, c2 = t0 >> 16;
    slot_cur = slot_next;
    if (c3 >= (uint32_t)kTopK) {
      lo = p3;
      c_lo = c3;
    } else if (c2 >= (uint32_t)kTopK) {
      lo = p2;
      c_lo = c2;
      hi = p3;
      c_hi = c3;
    } else if (c1 >= (uint32_t)kTopK) {
      lo = p1;
      c_lo = c1;
      hi = p2;
      c_hi = c2;
    } else {
      hi = p1;
      c_hi = c1;
    }
  }
  if (hi - lo > 1u) {
    // at most 32 keys lie in [lo, hi): collect them, rank them in warp 0
    const uint32_t span = hi - lo;
#pragma unroll
    for (int e = 0; e < 16; ++e)
      if (key[e] - lo < span) s_small[atomicAdd(&s_count[0], 1u)] = key[e];
    __syncthreads();
    if (warp == 0) {
      const uint32_t m = c_lo - c_hi, want_in = (uint32_t)kTopK - c_hi;   // m == s_count[0]
      const uint32_t k = lane < m ? s_small[lane] : 0u;
      uint32_t gt = 0, eq = 0;
#pragma unroll
      for (int j = 0; j < 32; ++j) {
        const uint32_t o = __shfl_sync(0xffffffffu, k, j);
        gt += (j < (int)m) && (o > k);
        eq += (j < (int)m) && (o == k);
      }
      const uint32_t hit = __ballot_sync(0xffffffffu, lane < m && gt < want_in && gt + eq >= want_in);
      if (lane == __ffs(hit) - 1) {
        s_sel[0] = k;
        s_sel[1] = c_hi + gt;
        s_sel[2] = eq;
      }
    }
    __syncthreads();
    lo = s_sel[0];
    c_hi = s_sel[1];
    c_lo = c_hi + s_sel[2];
  }
  // every key in [lo, lo + 1) equals lo
  const uint32_t thr = lo;
  const uint32_t want = (uint32_t)kTopK - c_hi;
  const uint32_t n_eq = c_lo - c_hi;

  // thr = the 200th largest magnitude; `want` of the n_eq elements equal to it are taken

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
    __syncthreads();   // red[] is reused below
  }

  // 6. tri-state bits (rawfingerprints.c:61-74); thread tid owns word tid of the 8192-bit array.
  // The reference compares (double)c with 0.001, -0.001 and 1.0.  0.001 lies strictly between two adjacent
  // floats, 0x3A83126E < 0.001 < 0x3A83126F = 0.001f, so (double)c > 0.001 <=> c >= 0.001f, and on the
  // magnitude bits that is an integer comparison; likewise |c| > 1.0 <=> key > 0x3F800000.  A selected
  // element (key >= thr) sets a bit iff key >= max(thr, bits(0.001f)): positive c by a signed comparison of
  // the float's bits, negative c by an unsigned one with the sign bit set.
  constexpr uint32_t kEpsBits = 0x3A83126Fu, kOneBits = 0x3F800000u;
  uint32_t word = 0, strong = 0;
  if (n_eq == want) {   // block-uniform, the usual case: every element equal to the threshold is selected
    const uint32_t t_bit = max(thr, kEpsBits), t_strong = max(thr, kOneBits + 1u);
    const int t_pos = (int)t_bit;
    const uint32_t t_neg = t_bit | 0x80000000u;
#pragma unroll
    for (int e = 0; e < 16; ++e) {   // three compares, each with one predicated update (spelled out: the compiler's
                                     // version went through SEL/LOP3 chains, twice the instructions)
      asm("{\n\t.reg .pred p, q, r;\n\t"
          "setp.ge.s32 p, %2, %4;\n\t"
          "setp.ge.u32 q, %2, %5;\n\t"
          "setp.ge.u32 r, %3, %6;\n\t"
          "@p or.b32 %0, %0, %7;\n\t"
          "@q or.b32 %0, %0, %8;\n\t"
          "@r add.u32 %1, %1, 1;\n\t}"
          : "+r"(word), "+r"(strong)
          : "r"(__float_as_uint(c[e])), "r"(key[e]), "r"(t_pos), "r"(t_neg), "r"(t_strong), "r"(1u << (2 * e)),
            "r"(2u << (2 * e)));
    }
  } else {
#pragma unroll
    for (int e = 0; e < 16; ++e) {
      bool sel = key[e] > thr;
      if (key[e] == thr) {
        sel = eq_before < want;
        ++eq_before;
      }
      if (sel) {
        if (key[e] >= kEpsBits) word |= 1u << (2 * e + (__float_as_uint(c[e]) >> 31));
        strong += key[e] > kOneBits;
      }
    }
  }
  // running count of set bits (warp scan now, cross-warp after the barrier): their positions feed MinHash
  const uint32_t nbits_mine = __popc(word);
  uint32_t nbits_incl = nbits_mine;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    const uint32_t y = __shfl_up_sync(0xffffffffu, nbits_incl, o);
    if (lane >= o) nbits_incl += y;
  }
  if (lane == 31) red[warp] = nbits_incl;
  strong = __reduce_add_sync(0xffffffffu, strong);
  if (lane == 0 && strong) atomicAdd(&s_strong, strong);
  __syncthreads();
  const bool silent = s_strong < 10u;   // rawfingerprints.c:96
  if (TAPS && tap_raw) reinterpret_cast<uint32_t*>(tap_raw + img * kRawBytes)[tid] = word;
  if (silent) {
    if (tid == 0) keep[img] = 0;
    return;
  }

  // 7. MinHash (minhash.c:13-28), turned inside out: instead of walking each of the 100 permutations until
  // it meets a set bit, every set bit b (at most 200) contributes its position in each permutation,
  // inv[b][p] (255 if b is not among the first 255 entries of permutation p), and the signature is the
  // element-wise minimum.  Row b of the inverse table is 100 x u16 = 25 x 8 bytes: 25 lanes take 4 permutations
  // each (two-lane 16-bit SIMD minima; sm_100 has no native 4 x 8-bit minimum), the 8 warps split the set
  // bits, and a last step combines the 8 partial minima.
  {
    uint32_t off = nbits_incl - nbits_mine, n_bits = 0;
#pragma unroll
    for (int w = 0; w < 8; ++w) {
      off += w < warp ? red[w] : 0u;
      n_bits += red[w];
    }
    uint32_t rest = word;
    while (rest) {   // list entry = index of the bit's row in 8-byte units (rows of the inverse table are 25 x 8 bytes)
      s_list[off++] = 25u * (32u * tid + __ffs(rest) - 1u);
      rest &= rest - 1;
    }
    // pad to a multiple of 4 with the all-255 row that follows the table
    const uint32_t n_chunks = (n_bits + 3u) >> 2;
    if (tid < 4 && n_bits + tid < 4u * n_chunks) s_list[n_bits + tid] = 25u * 8192u;
    __syncthreads();
    uint32_t m0 = 0xFFFFFFFFu, m1 = 0xFFFFFFFFu;   // permutations 4*lane, 4*lane+1 | 4*lane+2, 4*lane+3
    if (lane < 25) {
      const uint2* __restrict__ inv64 = reinterpret_cast<const uint2*>(inv);
      const uint32_t ul = (uint32_t)lane;   // 32-bit index arithmetic: one IMAD.WIDE per address
      for (uint32_t ch = warp; ch < n_chunks; ch += 8) {   // 4 bits per step: one 128-bit list read, 4 loads in flight
        const uint4 b = *reinterpret_cast<const uint4*>(&s_list[4 * ch]);
        const uint2 a0 = __ldg(inv64 + (b.x + ul)), a1 = __ldg(inv64 + (b.y + ul)), a2 = __ldg(inv64 + (b.z + ul)),
                    a3 = __ldg(inv64 + (b.w + ul));
        m0 = min3_u16x2(min3_u16x2(m0, a0.x, a1.x), a2.x, a3.x);
        m1 = min3_u16x2(min3_u16x2(m1, a0.y, a1.y), a2.y, a3.y);
      }
    }
    s_part[warp][lane] = make_uint2(m0, m1);
    __syncthreads();
    if (tid < 32) {
      uint32_t r0 = 0xFFFFFFFFu, r1 = 0xFFFFFFFFu;
#pragma unroll
      for (int w = 0; w < 8; w += 2) {
        const uint2 x = s_part[w][tid], y = s_part[w + 1][tid];
        r0 = min3_u16x2(r0, x.x, y.x);
        r1 = min3_u16x2(r1, x.y, y.y);
      }
      const uint32_t r = __byte_perm(r0, r1, 0x6420);   // four values <= 255 -> four bytes
      if (tid < 25) reinterpret_cast<uint32_t*>(sig_dense + img * kSigLen)[tid] = r;
      const uint32_t any = __ballot_sync(0xffffffffu, tid < 25 && r != 0xFFFFFFFFu);
      if (tid == 0) keep[img] = any ? 1 : 0;   // minhash.c:45: dropped when all 100 values are 255
    }
  }
}

void launch_k2_fingerprint(const TrackDev* tracks, const uint64_t* img_prefix, uint32_t n_tracks, uint64_t n_images,
                           const uint16_t* inv, const float* imax, const uint8_t* slot, const float* rec,
                           const float* tap_px, uint8_t* sig_dense, uint8_t* keep, uint8_t* tap_raw, float* tap_img,
                           float* tap_haar, cudaStream_t st) {
  if (n_images == 0) return;
  // the kernel itself no longer evaluates logf (the group records carry the scaled pixels' transforms)
  if (tap_raw || tap_img || tap_haar)
    k2_fingerprint_kernel<true><<<(unsigned)n_images, kK2Threads, 0, st>>>(
        tracks, img_prefix, n_tracks, inv, imax, slot, rec, tap_px, sig_dense, keep, tap_raw, tap_img, tap_haar);
  else
    k2_fingerprint_kernel<false><<<(unsigned)n_images, kK2Threads, 0, st>>>(
        tracks, img_prefix, n_tracks, inv, imax, slot, rec, tap_px, sig_dense, keep, tap_raw, tap_img, tap_haar);
}

// ---- device self-test of the logf mirror: compare against host values for a bit range ----
__global__ void logf_range_kernel(uint32_t lo_bits, uint32_t n, int variant, float* __restrict__ out) {
  __shared__ __align__(128) double s_logtab[32];   // same shared-memory table path as the fingerprint kernel
  if (threadIdx.x < 32) s_logtab[threadIdx.x] = c_logf_tab[threadIdx.x & 15][threadIdx.x >> 4];
  __syncthreads();
  const uint32_t saddr = (uint32_t)__cvta_generic_to_shared(s_logtab);
  const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const float x = __uint_as_float(lo_bits + i);
  out[i] = variant ? logf_glibc<1>(x, saddr) : logf_glibc<0>(x, saddr);
}

// ---- device self-test of the reciprocal divisions against IEEE division ----
// For every float x = bits lo..lo+n-1:
//   which 0: x / M_SQRT2 in double by reciprocal (one-step Markstein)        vs (float)((double)x / M_SQRT2)
//   which 1: x / 32767.0 in FP32 (div32767_f32, also the halving form)       vs (float)((double)x / 32767.0)
//   which 3: x / M_SQRT2 in FP32 (div_sqrt2_f32), x == 0 or |x| >= 2^-103    vs (float)((double)x / M_SQRT2)
//   which 4: x / c in FP32 (div_f32_recip), x == 0 or |x| >= 2^-106          vs x / c in IEEE f32, c = logf(256.0f)
// which 2: (255*v) / max for pseudo-random float pairs 0 <= v <= max (two-step Markstein, double results).
__global__ void div_test_kernel(int which, uint32_t lo_bits, uint64_t n, unsigned long long* __restrict__ bad, float c,
                                float rc) {
  const uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  double got, want;
  if (which != 2) {
    const float x = __uint_as_float((uint32_t)(lo_bits + i));
    if (!(fabsf(x) <= 3.0e38f) || __float_as_uint(x) == 0x80000000u) return;   // NaN / inf / -0 never reach the division
    float g, w;
    if (which == 0) {
      g = div_const_f32(x, kSqrt2, kInvSqrt2);
      w = __double2float_rn(__ddiv_rn((double)x, kSqrt2));
    } else if (which == 1) {
      g = div32767_f32<0>(x);
      w = __double2float_rn(__ddiv_rn((double)x, 32767.0));
      // the stereo form takes the sum before halving; (float)sum * 0.5f is exact unless it underflows
      const float g2 = div32767_f32<1>(x);
      const float w2 = __double2float_rn(__ddiv_rn((double)__fmul_rn(x, 0.5f), 32767.0));
      if (fabsf(x) >= 0x1p-100f && __float_as_uint(g2) != __float_as_uint(w2)) g = __uint_as_float(~__float_as_uint(w));
    } else if (which == 3) {
      if (x != 0.0f && fabsf(x) < kDivSqrt2MinAbs) return;
      g = div_sqrt2_f32(x);
      w = __double2float_rn(__ddiv_rn((double)x, kSqrt2));
    } else {
      if (x != 0.0f && fabsf(x) < kDivLog256MinAbs) return;
      g = div_f32_recip(x, c, rc);
      w = __fdiv_rn(x, c);
    }
    if (__float_as_uint(g) != __float_as_uint(w)) {
      atomicAdd(bad, 1ull);
      atomicMin(bad + 1, (unsigned long long)(uint32_t)(lo_bits + i));
    }
    return;
  }
  uint64_t z = (i + lo_bits) * 0x9E3779B97F4A7C15ull;
  z ^= z >> 29; z *= 0xBF58476D1CE4E5B9ull; z ^= z >> 32;
  uint32_t mb = 0x00800000u + (uint32_t)(z % (0x7f000000u - 0x00800000u));   // positive normal float
  z = z * 0x94D049BB133111EBull + 1; z ^= z >> 31;
  const float mx = __uint_as_float(mb);
  const uint32_t vb = (uint32_t)(z % ((uint64_t)mb + 1));                      // 0 <= v <= max (bit order)
  const double a = __dmul_rn(255.0, (double)__uint_as_float(vb)), b = (double)mx;
  got = div_markstein2(a, b, __drcp_rn(b));
  want = __ddiv_rn(a, b);
  if (__double_as_longlong(got) != __double_as_longlong(want)) {
    atomicAdd(bad, 1ull);
    atomicMin(bad + 1, (unsigned long long)i);
  }
}

void launch_div_test(int which, uint32_t lo_bits, uint64_t n, unsigned long long* bad, float c, float rc,
                     cudaStream_t st) {
  if (n) div_test_kernel<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(which, lo_bits, n, bad, c, rc);
}

void launch_logf_range(uint32_t lo_bits, uint32_t n, int variant, float* out, cudaStream_t st) {
  if (n) logf_range_kernel<<<(n + 255) / 256, 256, 0, st>>>(lo_bits, n, variant, out);
}

}  // namespace mnemo
