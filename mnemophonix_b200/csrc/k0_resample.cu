// k0_resample.cu — int16 PCM -> mono float -> 31-tap low-pass, decimate by 8.
//
// Replaces the downmix loop of read_samples (wav.c:358-375) and resample (resample.c:38-62).
// One CTA produces kTileOut consecutive 5512 Hz samples of one track: it stages the
// 8*kTileOut + 23 PCM frames it needs with 128-bit loads, down-mixes them into shared memory
// once (int sum -> f32 divide by the channel count -> divide by 32767 as the reference's f64 division rounds it), then every
// thread produces 4 consecutive outputs from 55 samples held in registers, running the
// reference's sequential mul-then-add tap loop for each (4 independent chains).
// Shared memory is padded by one word per 32 so that both the staging stores and the
// stride-32 register loads are bank-conflict-free.
#include "common.cuh"

namespace mnemo {

constexpr int kK0Threads = 256;
constexpr int kK0OutPerThread = 4;
constexpr int kTileOut = kK0Threads * kK0OutPerThread;          // 1024 outputs
constexpr int kTileIn = kTileOut * kDecim + (kTaps - kDecim);   // 8215 input frames
constexpr int kPerThreadIn = kK0OutPerThread * kDecim + (kTaps - kDecim);   // 55 samples
__device__ __forceinline__ int pad32(int i) { return i + (i >> 5); }
constexpr int kTileInPadded = kTileIn + (kTileIn >> 5) + 1;

int k0_tile_outputs() { return kTileOut; }

__constant__ float c_taps[kTaps];
void k0_upload_constants(const Tables& t) { cudaMemcpyToSymbol(c_taps, t.taps, sizeof(float) * kTaps); }

// wav.c:364-374: int sum; (float)sum / (float)channels in f32; then / 32767.0 in f64, back to f32.
// The f64 division is done in FP32 (div32767_f32, exact for every float, common.cuh); for two channels the
// exact halving is folded into it.
template <int CH>
__device__ __forceinline__ float downmix_value(float sum, float chf) {
  if (CH == 1) return div32767_f32<0>(sum);              // x / 1.0f
  if (CH == 2) return div32767_f32<1>(sum);              // x / 2.0f is exact
  return div32767_f32<0>(__fdiv_rn(sum, chf));
}
// (float)(a.lo * b0 + a.hi * b1) for the two int16 halves of a word and byte weights b (0 or 1), without an
// I2F: the dot product lands on the mantissa of 1.5 * 2^23 and a float subtraction removes the bias (exact,
// |sum| <= 2^16).
constexpr int kMagicBits = 0x4B400000;                   // 12582912.0f
__device__ __forceinline__ float pick_sum(int w, int weights) {
  return __fsub_rn(__int_as_float(__dp2a_lo(w, weights, kMagicBits)), 12582912.0f);
}

// INTERIOR: the whole tile lies inside the track, no per-frame bounds checks
template <int CH, bool INTERIOR>
__device__ __forceinline__ void stage_tile(float* __restrict__ mono, const TrackDev& tr, uint64_t in0) {
  const int16_t* __restrict__ pcm = tr.pcm;
  // frames past the end of the track contribute +0: acc + 0*h == acc, which is what the
  // reference's "start + j < n_samples" loop bound (resample.c:40) amounts to.
  if (CH == 1) {
    const int4* __restrict__ v = reinterpret_cast<const int4*>(pcm + in0);   // 16-byte aligned: in0 % 8 == 0
    constexpr int kVec = (kTileIn + 7) / 8;
    constexpr int kRounds = (kVec + kK0Threads - 1) / kK0Threads;   // 5
    int4 q[kRounds];
#pragma unroll
    for (int r = 0; r < kRounds; ++r) {   // loads first (see the two-channel case)
      const int i = threadIdx.x + r * kK0Threads;
      q[r] = make_int4(0, 0, 0, 0);
      if (i < kVec && (INTERIOR || in0 + (uint64_t)i * 8 < tr.n_pcm)) q[r] = __ldg(v + i);   // buffers carry >= 64 bytes of slack past n_pcm
    }
#pragma unroll
    for (int r = 0; r < kRounds; ++r) {
      const int i = threadIdx.x + r * kK0Threads;
      if (i >= kVec) break;
      const uint64_t f = in0 + (uint64_t)i * 8;
      const int w[4] = {q[r].x, q[r].y, q[r].z, q[r].w};
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        const float lo = pick_sum(w[k], 0x0001), hi = pick_sum(w[k], 0x0100);
        const int idx = i * 8 + 2 * k;
        if (idx < kTileIn) mono[pad32(idx)] = (INTERIOR || f + 2 * k < tr.n_pcm) ? downmix_value<1>(lo, 1.0f) : 0.0f;
        if (idx + 1 < kTileIn)
          mono[pad32(idx + 1)] = (INTERIOR || f + 2 * k + 1 < tr.n_pcm) ? downmix_value<1>(hi, 1.0f) : 0.0f;
      }
    }
  } else if (CH == 2) {
    const int4* __restrict__ v = reinterpret_cast<const int4*>(pcm + in0 * 2);
    constexpr int kVec = (kTileIn + 3) / 4;                       // 16-byte loads per tile
    constexpr int kRounds = (kVec + kK0Threads - 1) / kK0Threads;   // 9
    // all of a thread's loads are issued before the first one is used: the kernel is bound by DRAM latency
    // (one 16-byte load in flight per thread covered only ~2/3 of the bandwidth-delay product)
    int4 q[kRounds];
#pragma unroll
    for (int r = 0; r < kRounds; ++r) {
      const int i = threadIdx.x + r * kK0Threads;
      q[r] = make_int4(0, 0, 0, 0);
      if (i < kVec && (INTERIOR || in0 + (uint64_t)i * 4 < tr.n_pcm)) q[r] = __ldg(v + i);
    }
#pragma unroll
    for (int r = 0; r < kRounds; ++r) {
      const int i = threadIdx.x + r * kK0Threads;
      if (i >= kVec) break;
      const uint64_t f = in0 + (uint64_t)i * 4;
      const int w[4] = {q[r].x, q[r].y, q[r].z, q[r].w};
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        const float sum = pick_sum(w[k], 0x0101);
        const int idx = i * 4 + k;
        if (idx < kTileIn) mono[pad32(idx)] = (INTERIOR || f + k < tr.n_pcm) ? downmix_value<2>(sum, 2.0f) : 0.0f;
      }
    }
  } else {
    const uint32_t ch = tr.channels;
    const float chf = (float)ch;
    for (int i = threadIdx.x; i < kTileIn; i += kK0Threads) {
      const uint64_t f = in0 + i;
      float m = 0.0f;
      if (INTERIOR || f < tr.n_pcm) {
        int sum = 0;
        for (uint32_t c = 0; c < ch; ++c) sum += pcm[f * ch + c];
        m = downmix_value<0>((float)sum, chf);
      }
      mono[pad32(i)] = m;
    }
  }
}

__global__ void __launch_bounds__(kK0Threads) k0_resample_kernel(const TrackDev* __restrict__ tracks,
                                                                 const uint32_t* __restrict__ tile_prefix,
                                                                 uint32_t n_tracks, float* __restrict__ s5512) {
  __shared__ float mono[kTileInPadded];
  const uint32_t t = find_segment<uint32_t>(tile_prefix, n_tracks, blockIdx.x);
  const TrackDev tr = tracks[t];
  const uint32_t tile = blockIdx.x - tile_prefix[t];
  const uint64_t out0 = (uint64_t)tile * kTileOut;      // first output of this tile
  const uint64_t in0 = out0 * kDecim;                   // first PCM frame
  const bool interior = in0 + kTileIn + 8 <= tr.n_pcm;   // block-uniform
  if (tr.channels == 1) {
    if (interior) stage_tile<1, true>(mono, tr, in0);
    else stage_tile<1, false>(mono, tr, in0);
  } else if (tr.channels == 2) {
    if (interior) stage_tile<2, true>(mono, tr, in0);
    else stage_tile<2, false>(mono, tr, in0);
  } else {
    if (interior) stage_tile<0, true>(mono, tr, in0);
    else stage_tile<0, false>(mono, tr, in0);
  }
  __syncthreads();

  // outputs 4*tid .. 4*tid+3 read samples 32*tid .. 32*tid+54 (padded index 33*tid + j + (j >= 32))
  float x[kPerThreadIn];
  const float* base = mono + 33 * threadIdx.x;
#pragma unroll
  for (int j = 0; j < kPerThreadIn; ++j) x[j] = base[j + (j >> 5)];
  float acc[kK0OutPerThread];
#pragma unroll
  for (int m = 0; m < kK0OutPerThread; ++m) acc[m] = 0.0f;
#pragma unroll
  for (int j = 0; j < kTaps; ++j)   // resample.c:39-43: res += in[start + j] * h[j], j ascending
#pragma unroll
    for (int m = 0; m < kK0OutPerThread; ++m) acc[m] = __fadd_rn(acc[m], __fmul_rn(x[m * kDecim + j], c_taps[j]));

  float* __restrict__ out = s5512 + tr.s_off;
  const uint64_t oi = out0 + (uint64_t)threadIdx.x * kK0OutPerThread;
  if (oi + kK0OutPerThread <= tr.n5512) {
    *reinterpret_cast<float4*>(out + oi) = make_float4(acc[0], acc[1], acc[2], acc[3]);   // s_off, oi multiples of 4
  } else {
#pragma unroll
    for (int m = 0; m < kK0OutPerThread; ++m)
      if (oi + m < tr.n5512) out[oi + m] = acc[m];
  }
}

void launch_k0_resample(const TrackDev* tracks, const uint32_t* tile_prefix, uint32_t n_tracks, uint32_t n_tiles,
                        const Tables* tab, float* s5512, cudaStream_t st) {
  (void)tab;
  if (n_tiles == 0) return;
  k0_resample_kernel<<<n_tiles, kK0Threads, 0, st>>>(tracks, tile_prefix, n_tracks, s5512);
}

}  // namespace mnemo
