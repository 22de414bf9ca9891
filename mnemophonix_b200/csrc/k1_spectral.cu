// k1_spectral.cu — normalised 5512 Hz samples -> Hann -> 2048-point FFT -> 32 log-band energies.
//
// Replaces normalize()'s second loop (audionormalizer.c:22-31), frames_to_bins
// (spectralimages.c:80-107), fft/inplace_fft (fft.c:34-86) and calculate_bins (logbins.c:58-76).
//
// One warp transforms one frame entirely on-chip.  The reference's FFT is a radix-2
// decimation-in-time butterfly network on bit-reversed input whose every multiply and add is
// rounded separately; its results depend only on that dataflow graph, so the network is kept
// and only its schedule is changed:
//   * position p = 64*g + i (g = lane, i = register index): stages 1..6 (butterfly span < 64)
//     are 32 independent 64-point sub-transforms, one per lane, fully unrolled in registers
//     with warp-uniform twiddles read as constant-bank operands;
//   * one transposition through padded shared memory (bank-conflict-free both ways);
//   * complex values are 64-bit register pairs and the butterflies use Blackwell's packed FMUL2/FADD2
//     (two separately rounded f32 operations per issue slot): 6 instructions per butterfly instead of 10;
//   * stages 7..11 are 64 independent 32-point column transforms (fixed i, g = 0..31), two per
//     lane, again in registers, with per-lane twiddles in a lane-major shared table;
//   * the input is real, so positions 0 and L/2 of every sub-transform stay real: those
//     butterflies skip the multiplications by an exactly-zero imaginary part (values identical,
//     only the sign of a zero can differ, which no later operation can observe);
//   * only FFT outputs 118..742 feed the bands (logbins.c:64-75 with the edges of
//     logbins.c:44-55), so stage 11 computes the "a" half of 11 of its 16 butterflies per
//     column and dead-code elimination prunes what feeds nothing.
// The sequential per-band accumulation (logbins.c:68-73) is done by one lane per band.
//
// Input staging: a small elementwise kernel applies the normaliser's second loop
// (audionormalizer.c:22-31: s / rms, clipped to [-1, 1]) once per sample; the FFT kernel then pulls
// its 96-frame tiles (8128 samples, 32.5 KB) with TMA bulk copies (cp.async.bulk + mbarrier) into a
// double buffer, so the next tile streams in while the current one is transformed.
#include <math.h>

#include "common.cuh"

namespace mnemo {

constexpr int kK1Warps = 12;
constexpr int kK1Threads = kK1Warps * 32;
constexpr int kK1FramesPerWarp = 8;
constexpr int kK1TileFrames = kK1Warps * kK1FramesPerWarp;            // 96 frames
constexpr int kK1TileSamples = kHop * (kK1TileFrames - 1) + kFrame;   // 8128 samples
constexpr int kScratchStride = 68;   // floats per row (32 complex + pad); 68 = 4 mod 32: STS.64 columns and LDS.128 rows conflict-free
constexpr int kScratchFloats = 32 * kScratchStride;                   // per warp (also holds the 704 power terms)
constexpr int kTwHiFloat2 = 2 * 31 * 32;
constexpr int kMaxBandWidth = 44;   // >= the widest band (41 FFT bins with the reference's edges; checked at context creation)
int k1_max_band_width() { return kMaxBandWidth; }
constexpr size_t kK1Smem =
    sizeof(float) * (2 * kK1TileSamples + kFrame + kK1Warps * kScratchFloats) + sizeof(float2) * kTwHiFloat2 + 16;

int k1_tile_frames() { return kK1TileFrames; }

__constant__ float2 c_tw64[32];
__constant__ float c_one;   // 1.0f at run time (see twiddle_mul)

void k1_upload_constants(const Tables& t) {
  const float one = 1.0f;
  cudaMemcpyToSymbol(c_tw64, t.tw64, sizeof(float2) * 32);
  cudaMemcpyToSymbol(c_one, &one, sizeof(float));
}

__host__ __device__ constexpr int rev6(int v) {
  return ((v & 1) << 5) | ((v & 2) << 3) | ((v & 4) << 1) | ((v & 8) >> 1) | ((v & 16) >> 3) | ((v & 32) >> 5);
}

// ---- packed complex arithmetic (Blackwell f32x2: FMUL2 / FADD2, two IEEE-rounded f32 ops per issue) ----
// A complex value is a 64-bit register pair (re, im).  Only mul and add/sub are ever used packed,
// and never a packed mul feeding a packed add: ptxas contracts mul.rn.f32x2 + add.rn.f32x2 into
// FFMA2 even with -fmad=false, which would change the rounding.  The products are therefore
// combined by two scalar adds, exactly as the reference's "w_re*re - w_im*im" / "w_re*im + w_im*re".
typedef uint64_t cplx;
__device__ __forceinline__ cplx pk(float re, float im) {
  cplx r;
  asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(re), "f"(im));
  return r;
}
__device__ __forceinline__ void upk(cplx v, float& re, float& im) { asm("mov.b64 {%0, %1}, %2;" : "=f"(re), "=f"(im) : "l"(v)); }
__device__ __forceinline__ float lo(cplx v) {
  float a, b;
  upk(v, a, b);
  return a;
}
__device__ __forceinline__ float hi(cplx v) {
  float a, b;
  upk(v, a, b);
  return b;
}
__device__ __forceinline__ cplx mul2(cplx a, cplx b) {
  cplx r;
  asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b));
  return r;
}
__device__ __forceinline__ cplx add2(cplx a, cplx b) {
  cplx r;
  asm("add.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b));
  return r;
}
__device__ __forceinline__ cplx sub2(cplx a, cplx b) {
  cplx r;
  asm("sub.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b));
  return r;
}

// tao = w * b with every product and sum rounded on its own (fft.c:53-54)
// (wr*br - wi*bi, wr*bi + wi*br) with all four products and both sums rounded on their own (fft.c:53-54), and no
// scalar FP32 instruction.  The second product is taken on the swapped operand with the sign folded into the
// twiddle, q' = (bi * -wi, br * wi) — half-swap and per-half negation are free operand modifiers of the packed
// instructions — and the sum is an FMA by one, fma(p, 1, q') = RN(p + q').  The 1 must be a RUN-TIME value
// (c_one, set by the host): a packed add of a packed product, or an FMA by a literal 1, is contracted by ptxas
// into an FFMA2 that rounds the product away, -fmad=false or not.  (Scalar FP32 instructions between packed ones
// cost up to two cycles each, tools/fp32x2_probe.cu; with them this kernel ran at 2.16 ms instead of 1.9.)
__device__ __forceinline__ cplx twiddle_mul(cplx b, const float wr, const float wi) {
  const cplx p = mul2(b, pk(wr, wr));
  const cplx qs = mul2(pk(hi(b), lo(b)), pk(-wi, wi));
  const float one = c_one;
  cplx t;
  asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(t) : "l"(p), "l"(pk(one, one)), "l"(qs));
  return t;
}

// fft.c:53-62: 2 FMUL2 + 1 FFMA2 + 2 FADD2
__device__ __forceinline__ void bfly(cplx& a, cplx& b, const float wr, const float wi) {
  const cplx t = twiddle_mul(b, wr, wi);
  b = sub2(a, t);
  a = add2(a, t);
}

// Stage S (span L = 2^S <= 64) of the 64-point sub-transform held in registers.
template <int S>
__device__ __forceinline__ void stage_lo(cplx (&z)[64]) {
  constexpr int L = 1 << S, H = L >> 1;
#pragma unroll
  for (int base = 0; base < 64; base += L) {
#pragma unroll
    for (int k = 0; k < H; ++k) {
      const int a = base + k, b = a + H;
      const int m = k * (64 / L);   // twiddle (l = L, k) == W64^m
      if (k == 0) {
        // both inputs real, w = (1, -0): tao == b exactly; the imaginary halves stay (dead) zeros
        const float x = lo(z[a]), y = lo(z[b]);
        z[a] = pk(__fadd_rn(x, y), 0.0f);
        z[b] = pk(__fsub_rn(x, y), 0.0f);
      } else if (2 * k == H) {
        // both inputs real: tao = (wr*b, wi*b); outputs (a + tr, 0 + ti), (a - tr, 0 - ti)
        const float x = lo(z[a]), y = lo(z[b]);
        const float tr = __fmul_rn(c_tw64[m].x, y), ti = __fmul_rn(c_tw64[m].y, y);
        z[b] = pk(__fsub_rn(x, tr), -ti);
        z[a] = pk(__fadd_rn(x, tr), ti);
      } else {
        bfly(z[a], z[b], c_tw64[m].x, c_tw64[m].y);
      }
    }
  }
}

// Stages 7..11 on one column (fixed i; element index = g) and the band-power terms.
// tw points at tw_hi[h][0][lane]; consecutive slots are 32 float2 apart.
__device__ __forceinline__ void column_pass(cplx (&c)[32], const float2* __restrict__ tw, float* __restrict__ pw_col) {
  {  // stage 7: span 128, k = i
    const float2 w = tw[0];
#pragma unroll
    for (int g = 0; g < 32; g += 2) bfly(c[g], c[g + 1], w.x, w.y);
  }
#pragma unroll
  for (int q = 0; q < 2; ++q) {  // stage 8: span 256, k = 64*(g&1) + i
    const float2 w = tw[32 * (1 + q)];
#pragma unroll
    for (int g = q; g < 32; g += 4) bfly(c[g], c[g + 2], w.x, w.y);
  }
#pragma unroll
  for (int q = 0; q < 4; ++q) {  // stage 9: span 512
    const float2 w = tw[32 * (3 + q)];
#pragma unroll
    for (int g = q; g < 32; g += 8) bfly(c[g], c[g + 4], w.x, w.y);
  }
#pragma unroll
  for (int q = 0; q < 8; ++q) {  // stage 10: span 1024
    const float2 w = tw[32 * (7 + q)];
#pragma unroll
    for (int g = q; g < 32; g += 16) bfly(c[g], c[g + 8], w.x, w.y);
  }
  // stage 11: span 2048; outputs 118..742 live in g = 1..11, "a" side only
#pragma unroll
  for (int g = 1; g <= 11; ++g) {
    const float2 w = tw[32 * (15 + g)];
    const cplx x = add2(c[g], twiddle_mul(c[g + 16], w.x, w.y));
    // logbins.c:69-71: re/1024.0 and im/1024.0 (exact scalings, rounded to f32), squares, sum
    const cplx r = mul2(x, pk(0.0009765625f, 0.0009765625f));
    const cplx sq = mul2(r, r);
    pw_col[64 * (g - 1)] = __fadd_rn(lo(sq), hi(sq));
  }
}

// ---- mbarrier / TMA bulk copy (1-D) ----
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "WAIT_%=:\n"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
      "@p bra DONE_%=;\n"
      "bra WAIT_%=;\n"
      "DONE_%=:\n"
      "}\n" ::"r"(smem_u32(bar)),
      "r"(parity)
      : "memory");
}
__device__ __forceinline__ void tma_load_1d(void* dst_smem, const void* src_gmem, uint32_t bytes, uint64_t* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                   smem_u32(dst_smem)),
               "l"(src_gmem), "r"(bytes), "r"(smem_u32(bar))
               : "memory");
}

// audionormalizer.c:22-31, one block per 4096-sample chunk of a track
__global__ void __launch_bounds__(256) normalize_kernel(const TrackDev* __restrict__ tracks,
                                                        const uint32_t* __restrict__ chunk_track, uint64_t n_chunks,
                                                        const float* __restrict__ s5512, const float* __restrict__ rms,
                                                        float* __restrict__ out) {
  const uint64_t c = blockIdx.x;
  if (c >= n_chunks) return;
  const uint32_t t = chunk_track[c];
  const TrackDev tr = tracks[t];
  const uint32_t first = (uint32_t)(c - tr.chunk_off) * kSumChunk;
  const uint32_t cnt = min((uint32_t)kSumChunk, tr.n5512 - first);
  const float r = rms[t];
  const float4* __restrict__ src = reinterpret_cast<const float4*>(s5512 + tr.s_off + first);
  float4* __restrict__ dst = reinterpret_cast<float4*>(out + tr.s_off + first);
  auto nrm = [r](float x) {
    float v = __fdiv_rn(x, r);
    if (v < -1.0f) v = -1.0f;
    else if (v > 1.0f) v = 1.0f;
    return v;
  };
  // chunks start on 16-byte boundaries and buffers are padded to 64 floats: whole float4s
  for (uint32_t i = threadIdx.x; i < (cnt + 3) / 4; i += 256) {
    float4 q = src[i];
    q.x = nrm(q.x);
    q.y = nrm(q.y);
    q.z = nrm(q.z);
    q.w = nrm(q.w);
    dst[i] = q;
  }
}

__global__ void __launch_bounds__(kK1Threads, 1)
    k1_spectral_kernel(const TrackDev* __restrict__ tracks, const uint32_t* __restrict__ tile_prefix, uint32_t n_tracks,
                       uint32_t n_tiles, const Tables* __restrict__ tab, const float* __restrict__ snorm,
                       float* __restrict__ bins) {
  extern __shared__ __align__(128) float smem[];
  float* const tile0 = smem;                              // 2 x kK1TileSamples (TMA destinations)
  float* hann = smem + 2 * kK1TileSamples;                // kFrame
  float2* tw_hi = reinterpret_cast<float2*>(hann + kFrame);   // [2][31][32]
  float* scratch_all = reinterpret_cast<float*>(tw_hi + kTwHiFloat2);
  uint64_t* bars = reinterpret_cast<uint64_t*>(scratch_all + kK1Warps * kScratchFloats);

  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  float* sc = scratch_all + warp * kScratchFloats;

  for (int i = threadIdx.x; i < kFrame; i += kK1Threads) hann[i] = tab->hann[i];
  for (int i = threadIdx.x; i < kTwHiFloat2; i += kK1Threads) tw_hi[i] = (&tab->tw_hi[0][0][0])[i];
  const int e_lo = (int)tab->edges[lane] - 64, e_hi = (int)tab->edges[lane + 1] - 64;
  const float width = (float)(unsigned)(e_hi - e_lo);   // logbins.c:75
  const int rl = (int)(__brev((unsigned)lane) >> 27);   // rev5(lane)

  if (threadIdx.x == 0) {
    mbar_init(&bars[0], 1);
    mbar_init(&bars[1], 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();

  // tile -> (track, first frame, frame count); issue its TMA copy into buffer `buf`
  auto issue = [&](uint32_t tix, int buf) {
    const uint32_t t = find_segment<uint32_t>(tile_prefix, n_tracks, tix);
    const uint32_t f0 = (tix - tile_prefix[t]) * kK1TileFrames;
    const uint32_t nf = min((uint32_t)kK1TileFrames, tracks[t].n_frames - f0);
    const uint32_t bytes = (kHop * (nf - 1) + kFrame) * (uint32_t)sizeof(float);
    mbar_expect_tx(&bars[buf], bytes);
    tma_load_1d(tile0 + buf * kK1TileSamples, snorm + tracks[t].s_off + (uint64_t)f0 * kHop, bytes, &bars[buf]);
  };
  if (threadIdx.x == 0 && blockIdx.x < n_tiles) issue(blockIdx.x, 0);
  uint32_t phase_bits = 0;   // bit b = parity to wait for on bars[b]
  int cur = 0;

  for (uint32_t tix = blockIdx.x; tix < n_tiles; tix += gridDim.x, cur ^= 1) {
    const uint32_t t = find_segment<uint32_t>(tile_prefix, n_tracks, tix);
    const TrackDev tr = tracks[t];
    const uint32_t f0 = (tix - tile_prefix[t]) * kK1TileFrames;
    const uint32_t nf = min((uint32_t)kK1TileFrames, tr.n_frames - f0);
    // the other buffer was released by the __syncthreads that ended the previous iteration
    if (threadIdx.x == 0 && tix + gridDim.x < n_tiles) issue(tix + gridDim.x, cur ^ 1);
    mbar_wait(&bars[cur], (phase_bits >> cur) & 1u);
    phase_bits ^= 1u << cur;
    const float* tile = tile0 + cur * kK1TileSamples;

    for (uint32_t fl = warp; fl < nf; fl += kK1Warps) {
      cplx z[64];
      {  // spectralimages.c:94-98 Hann multiply, loaded straight into bit-reversed order (fft.c:35-45)
        const float* fs = tile + fl * kHop + rl;
        const float* hw = hann + rl;
#pragma unroll
        for (int r = 0; r < 64; ++r) z[rev6(r)] = pk(__fmul_rn(fs[32 * r], hw[32 * r]), 0.0f);
      }
      stage_lo<1>(z);
      stage_lo<2>(z);
      stage_lo<3>(z);
      stage_lo<4>(z);
      stage_lo<5>(z);
      stage_lo<6>(z);

      // transposition: lane g owns positions 64g+i; afterwards lane t owns i = t and t+32, all g.
      // Two rounds (i < 32, then i >= 32) through a 32-row scratch: row i holds the 32 complex values
      // of column i; 64-bit stores down a column, 128-bit loads along a row, both conflict-free.
      cplx cz[2][32];
#pragma unroll
      for (int h = 0; h < 2; ++h) {
        cplx* sc2 = reinterpret_cast<cplx*>(sc);
#pragma unroll
        for (int i = 0; i < 32; ++i) sc2[i * (kScratchStride / 2) + lane] = z[i + 32 * h];
        __syncwarp();
        const ulonglong2* row = reinterpret_cast<const ulonglong2*>(sc + lane * kScratchStride);
#pragma unroll
        for (int c = 0; c < 16; ++c) {
          const ulonglong2 q = row[c];
          cz[h][2 * c] = q.x;
          cz[h][2 * c + 1] = q.y;
        }
        __syncwarp();
      }

      // columns; the scratch now receives the power terms of outputs 64..767 (index p - 64)
      column_pass(cz[0], tw_hi + lane, sc + lane);
      column_pass(cz[1], tw_hi + 31 * 32 + lane, sc + 32 + lane);
      __syncwarp();

      // logbins.c:64-75: sequential sum over the band's FFT bins, one band per lane
      // Fully unrolled over the widest band (41 bins, App. B of SURVEY.md): one predicated load at an immediate
      // offset and one add per slot; "+ 0.0f" past the band end is a no-op (the sum is never -0).
      float sum = 0.0f;
      {
        const float* __restrict__ pb = sc + e_lo;
        const int w = e_hi - e_lo;
#pragma unroll
        for (int u0 = 0; u0 < kMaxBandWidth; u0 += 8) {
          float tpw[8];   // loads first, then the dependent adds
#pragma unroll
          for (int u = 0; u < 8; ++u) tpw[u] = (u0 + u < w) ? pb[u0 + u] : 0.0f;
#pragma unroll
          for (int u = 0; u < 8; ++u)
            if (u0 + u < kMaxBandWidth) sum = __fadd_rn(sum, tpw[u]);
        }
      }
      bins[(tr.frame_off + f0 + fl) * kBands + lane] = __fdiv_rn(sum, width);
      __syncwarp();
    }
    __syncthreads();   // every warp is done with tile_buf[cur]: it may be refilled next iteration
  }
}

// =====================================================================================================================
// Tolerance mode ("fast K1", selected per context: mnemo_create_ex(..., MNEMO_FAST_SPECTRA), never the default).
//
// BASELINE.json allows 1e-4 relative error on the spectral images; the kernel above spends its time replaying the
// reference's radix-2 network rounding for rounding.  This one computes the same band energies the cheap way:
//   * the frame is real, so its 2048-point spectrum comes from ONE 1024-point complex transform of
//     z[n] = x[2n] + i x[2n+1] and an untangling step, X[k] = ((Z[k] + conj Z[1024-k]) - i W2048^k (Z[k] - conj Z[1024-k])) / 2,
//     evaluated only for the 625 bins 118..742 that feed the bands (logbins.c:64-75);
//   * the 1024-point transform is a 32 x 32 four-step: 32-point transform per lane in registers (decimation in
//     frequency, twiddles as constant-bank operands), twiddle by W1024^(b*c), one transposition through shared
//     memory, second 32-point transform per lane; complex products are 1 FMUL2 + 1 FFMA2 (fused, unlike above);
//   * twiddles come from double-precision cos/sin rounded to float (more accurate than the reference's cosf/sinf of
//     a float angle);
//   * band sums are four-way parallel partial sums instead of the reference's sequential order.
// About 750 packed FP32 instructions per lane and frame against 1870 FMA-pipe instructions of the exact replay.
// The results differ from the reference's in the last bits of the band energies (measured gates: bench.py
// "fast_spectra" and tools/fast_gates.py); everything downstream (group records, K2) is unchanged.

constexpr int kF1Warps = 16;
constexpr int kF1Threads = kF1Warps * 32;
constexpr int kF1TileFrames = kK1TileFrames;                     // same tiling as the exact kernel (launch code shared)
constexpr int kF1TStride = 33;                                   // float2 per row of the transposition scratch
constexpr int kF1ScratchF2 = 32 * kF1TStride;                    // per warp
constexpr int kF1FirstD = 3, kF1LastD = 23;                      // k = c + 32 d covers 118..742 for d = 3..23
constexpr int kF1NumD = kF1LastD - kF1FirstD + 1;
constexpr size_t kF1Smem = sizeof(float) * (2 * kK1TileSamples + kFrame) + sizeof(float2) * (1024 + kF1NumD * 32) +
                           sizeof(float2) * kF1Warps * kF1ScratchF2 + 16;

__constant__ float2 c_w32[16];   // W32^m = exp(-2 pi i m / 32), m < 16

__device__ __forceinline__ cplx fma2(cplx a, cplx b, cplx c) {
  cplx r;
  asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(r) : "l"(a), "l"(b), "l"(c));
  return r;
}
// b * w, fused: (br wr - bi wi, bi wr + br wi)
__device__ __forceinline__ cplx cmul_f(cplx b, const float wr, const float wi) {
  return fma2(pk(hi(b), lo(b)), pk(-wi, wi), mul2(b, pk(wr, wr)));
}
// b * (-i) = (bi, -br)
__device__ __forceinline__ cplx mul_mi(cplx b) { return pk(hi(b), -lo(b)); }

__host__ __device__ constexpr int rev5c(int v) {
  return ((v & 1) << 4) | ((v & 2) << 2) | (v & 4) | ((v & 8) >> 2) | ((v & 16) >> 4);
}

// 32-point forward transform in registers, decimation in frequency: natural order in, out[k] = v[rev5c(k)].
__device__ __forceinline__ void fft32_dif(cplx (&v)[32]) {
#pragma unroll
  for (int s = 0; s < 5; ++s) {
    const int L = 32 >> s, H = L >> 1;
#pragma unroll
    for (int base = 0; base < 32; base += L) {
#pragma unroll
      for (int k = 0; k < H; ++k) {
        const int m = k * (32 / L);   // W_L^k = W32^m
        const cplx a = v[base + k], b = v[base + k + H];
        v[base + k] = add2(a, b);
        const cplx d = sub2(a, b);
        if (m == 0) v[base + k + H] = d;
        else if (m == 8) v[base + k + H] = mul_mi(d);
        else v[base + k + H] = cmul_f(d, c_w32[m].x, c_w32[m].y);
      }
    }
  }
}

__global__ void __launch_bounds__(kF1Threads, 1)
    k1_fast_kernel(const TrackDev* __restrict__ tracks, const uint32_t* __restrict__ tile_prefix, uint32_t n_tracks,
                   uint32_t n_tiles, const Tables* __restrict__ tab, const float2* __restrict__ tw_fast,
                   const float* __restrict__ snorm, float* __restrict__ bins) {
  extern __shared__ __align__(128) float smem[];
  float* const tile0 = smem;                                        // 2 x kK1TileSamples (TMA destinations)
  float* hann = smem + 2 * kK1TileSamples;                          // kFrame
  float2* tw4 = reinterpret_cast<float2*>(hann + kFrame);           // [c][b]: W1024^(b c)
  float2* twu = tw4 + 1024;                                         // [d - 3][c]: W2048^(c + 32 d)
  float2* scratch_all = twu + kF1NumD * 32;
  uint64_t* bars = reinterpret_cast<uint64_t*>(scratch_all + kF1Warps * kF1ScratchF2);

  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  float2* sc = scratch_all + warp * kF1ScratchF2;
  float* pw = reinterpret_cast<float*>(sc);   // after the transposition: power terms of bins 96..767 (index k - 96)

  for (int i = threadIdx.x; i < kFrame; i += kF1Threads) hann[i] = tab->hann[i];
  for (int i = threadIdx.x; i < 1024 + kF1NumD * 32; i += kF1Threads) tw4[i] = tw_fast[i];
  const int e_lo = (int)tab->edges[lane] - 96, e_hi = (int)tab->edges[lane + 1] - 96;
  const float width = (float)(unsigned)(e_hi - e_lo);

  if (threadIdx.x == 0) {
    mbar_init(&bars[0], 1);
    mbar_init(&bars[1], 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();

  auto issue = [&](uint32_t tix, int buf) {
    const uint32_t t = find_segment<uint32_t>(tile_prefix, n_tracks, tix);
    const uint32_t f0 = (tix - tile_prefix[t]) * kF1TileFrames;
    const uint32_t nf = min((uint32_t)kF1TileFrames, tracks[t].n_frames - f0);
    const uint32_t bytes = (kHop * (nf - 1) + kFrame) * (uint32_t)sizeof(float);
    mbar_expect_tx(&bars[buf], bytes);
    tma_load_1d(tile0 + buf * kK1TileSamples, snorm + tracks[t].s_off + (uint64_t)f0 * kHop, bytes, &bars[buf]);
  };
  if (threadIdx.x == 0 && blockIdx.x < n_tiles) issue(blockIdx.x, 0);
  uint32_t phase_bits = 0;
  int cur = 0;
  const int src_lane = (32 - lane) & 31;   // Z[1024 - k] lives in lane (32 - c) mod 32

  for (uint32_t tix = blockIdx.x; tix < n_tiles; tix += gridDim.x, cur ^= 1) {
    const uint32_t t = find_segment<uint32_t>(tile_prefix, n_tracks, tix);
    const TrackDev tr = tracks[t];
    const uint32_t f0 = (tix - tile_prefix[t]) * kF1TileFrames;
    const uint32_t nf = min((uint32_t)kF1TileFrames, tr.n_frames - f0);
    if (threadIdx.x == 0 && tix + gridDim.x < n_tiles) issue(tix + gridDim.x, cur ^ 1);
    mbar_wait(&bars[cur], (phase_bits >> cur) & 1u);
    phase_bits ^= 1u << cur;
    const float* tile = tile0 + cur * kK1TileSamples;

    for (uint32_t fl = warp; fl < nf; fl += kF1Warps) {
      cplx v[32];
      {  // z[32 a + b] = (x[2n] h[2n], x[2n+1] h[2n+1]), n = 32 a + b, b = lane
        const cplx* fs = reinterpret_cast<const cplx*>(tile + fl * kHop) + lane;   // tile + 64 fl is 8-byte aligned
        const cplx* hw = reinterpret_cast<const cplx*>(hann) + lane;
#pragma unroll
        for (int a = 0; a < 32; ++a) v[a] = mul2(fs[32 * a], hw[32 * a]);
      }
      fft32_dif(v);   // over a: Y_b[c] = v[rev5c(c)]
      // twiddle by W1024^(b c) and transpose: scratch[c][b]
#pragma unroll
      for (int c = 0; c < 32; ++c) {
        cplx y = v[rev5c(c)];
        if (c) {
          const float2 w = tw4[c * 32 + lane];
          y = cmul_f(y, w.x, w.y);
        }
        reinterpret_cast<cplx*>(sc)[c * kF1TStride + lane] = y;
      }
      __syncwarp();
#pragma unroll
      for (int b = 0; b < 32; ++b) v[b] = reinterpret_cast<const cplx*>(sc)[lane * kF1TStride + b];
      __syncwarp();
      fft32_dif(v);   // over b: Z[c + 32 d] = v[rev5c(d)], c = lane

      // untangle bins k = c + 32 d, d = 3..23, and leave |X[k] / 1024|^2 in the scratch at k - 96
#pragma unroll
      for (int d = kF1FirstD; d <= kF1LastD; ++d) {
        const cplx A = v[rev5c(d)];
        // partner Z[1024 - k]: register 31 - d of lane 32 - c (register 32 - d of lane 0 for c == 0)
        const cplx send = lane == 0 ? v[rev5c(32 - d)] : v[rev5c(31 - d)];
        float pr, pi;
        upk(send, pr, pi);
        pr = __shfl_sync(0xffffffffu, pr, src_lane);
        pi = __shfl_sync(0xffffffffu, pi, src_lane);
        const cplx Bc = pk(pr, -pi);                       // conj Z[1024 - k]
        const cplx S = add2(A, Bc), D = sub2(A, Bc);
        const float2 w = twu[(d - kF1FirstD) * 32 + lane];
        const cplx q = cmul_f(D, w.x, w.y);                // W2048^k (A - conj B)
        const cplx X2 = add2(S, mul_mi(q));                // 2 X[k]
        const cplx r = mul2(X2, pk(0.00048828125f, 0.00048828125f));   // X[k] / 1024
        const cplx sq = mul2(r, r);
        pw[32 * (d - kF1FirstD) + lane] = lo(sq) + hi(sq);
      }
      __syncwarp();
      // band means (logbins.c:64-75), one band per lane, four partial sums
      float s0 = 0.0f, s1 = 0.0f, s2 = 0.0f, s3 = 0.0f;
      {
        const float* __restrict__ pb = pw + e_lo;
        const int w = e_hi - e_lo;
#pragma unroll
        for (int u = 0; u < kMaxBandWidth; u += 4) {
          s0 += (u < w) ? pb[u] : 0.0f;
          s1 += (u + 1 < w) ? pb[u + 1] : 0.0f;
          s2 += (u + 2 < w) ? pb[u + 2] : 0.0f;
          s3 += (u + 3 < w) ? pb[u + 3] : 0.0f;
        }
      }
      bins[(tr.frame_off + f0 + fl) * kBands + lane] = __fdiv_rn((s0 + s1) + (s2 + s3), width);
      __syncwarp();
    }
    __syncthreads();
  }
}

// host side of the tolerance mode: the twiddle tables, from double precision
size_t k1_fast_table_float2() { return 1024 + (size_t)kF1NumD * 32; }
void k1_fast_tables(float2* tw /* k1_fast_table_float2() entries */, float2 w32[16]) {
  const double two_pi = 6.283185307179586476925286766559;
  for (int c = 0; c < 32; ++c)
    for (int b = 0; b < 32; ++b) {
      const double ang = -two_pi * (double)(b * c) / 1024.0;
      tw[c * 32 + b] = make_float2((float)cos(ang), (float)sin(ang));
    }
  for (int d = kF1FirstD; d <= kF1LastD; ++d)
    for (int c = 0; c < 32; ++c) {
      const double ang = -two_pi * (double)(c + 32 * d) / 2048.0;
      tw[1024 + (d - kF1FirstD) * 32 + c] = make_float2((float)cos(ang), (float)sin(ang));
    }
  for (int m = 0; m < 16; ++m) {
    const double ang = -two_pi * (double)m / 32.0;
    w32[m] = make_float2((float)cos(ang), (float)sin(ang));
  }
}
void k1_fast_upload_constants(const float2 w32[16]) { cudaMemcpyToSymbol(c_w32, w32, sizeof(float2) * 16); }

void launch_k1_fast(const TrackDev* tracks, const uint32_t* tile_prefix, uint32_t n_tracks, uint32_t n_tiles, const Tables* tab,
                    const float2* tw_fast, const float* snorm, float* bins, int sm_count, cudaStream_t st) {
  if (n_tiles == 0) return;
  cudaFuncSetAttribute(k1_fast_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kF1Smem);
  const unsigned grid = n_tiles < (uint32_t)sm_count ? n_tiles : (unsigned)sm_count;
  k1_fast_kernel<<<grid, kF1Threads, kF1Smem, st>>>(tracks, tile_prefix, n_tracks, n_tiles, tab, tw_fast, snorm, bins);
}

void launch_normalize(const TrackDev* tracks, const uint32_t* chunk_track, uint64_t n_chunks, const float* s5512,
                      const float* rms, float* snorm, cudaStream_t st) {
  if (n_chunks) normalize_kernel<<<(unsigned)n_chunks, 256, 0, st>>>(tracks, chunk_track, n_chunks, s5512, rms, snorm);
}

void launch_k1_spectral(const TrackDev* tracks, const uint32_t* tile_prefix, uint32_t n_tracks, uint32_t n_tiles,
                        const Tables* tab, const float* snorm, float* bins, int sm_count, cudaStream_t st) {
  if (n_tiles == 0) return;
  cudaFuncSetAttribute(k1_spectral_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kK1Smem);
  const unsigned grid = n_tiles < (uint32_t)sm_count ? n_tiles : (unsigned)sm_count;
  k1_spectral_kernel<<<grid, kK1Threads, kK1Smem, st>>>(tracks, tile_prefix, n_tracks, n_tiles, tab, snorm, bins);
}

}  // namespace mnemo
