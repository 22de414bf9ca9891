// k0_resample.cu — int16 PCM -> mono float -> 31-tap low-pass, decimate by 8.
//
// Replaces the downmix loop of read_samples (wav.c:358-375) and resample (resample.c:38-62).
// One CTA produces kTileOut consecutive 5512 Hz samples of one track: it stages the
// 8*kTileOut + 23 PCM frames it needs with 128-bit loads, down-mixes them into shared memory
// once (int sum -> f32 divide by the channel count -> f64 divide by 32767 -> f32), then every
// thread runs the reference's sequential mul-then-add tap loop on a bank-conflict-free view.
#include "common.cuh"

namespace mnemo {

constexpr int kK0Threads = 256;
constexpr int kK0OutPerThread = 4;
constexpr int kTileOut = kK0Threads * kK0OutPerThread;          // 1024 outputs
constexpr int kTileIn = kTileOut * kDecim + (kTaps - kDecim);   // 8215 input frames
// pad one word every 32 so the stride-8 tap reads hit 32 distinct banks
__device__ __forceinline__ int pad32(int i) { return i + (i >> 5); }
constexpr int kTileInPadded = kTileIn + (kTileIn >> 5) + 1;

int k0_tile_outputs() { return kTileOut; }

// wav.c:364-374: int sum; (float)sum / (float)channels in f32; then / 32767.0 in f64.
__device__ __forceinline__ float downmix_value(int sum, float chf) {
  float mean = __fdiv_rn((float)sum, chf);
  return div_const_f32(mean, 32767.0, kInv32767);   // (float)((double)mean / 32767.0), see common.cuh
}

__global__ void __launch_bounds__(kK0Threads) k0_resample_kernel(const TrackDev* __restrict__ tracks,
                                                                 const uint32_t* __restrict__ tile_prefix,
                                                                 uint32_t n_tracks, const Tables* __restrict__ tab,
                                                                 float* __restrict__ s5512) {
  __shared__ float mono[kTileInPadded];
  __shared__ float taps[kTaps];
  const uint32_t t = find_segment<uint32_t>(tile_prefix, n_tracks, blockIdx.x);
  const TrackDev tr = tracks[t];
  const uint32_t tile = blockIdx.x - tile_prefix[t];
  const uint64_t out0 = (uint64_t)tile * kTileOut;      // first output of this tile
  const uint64_t in0 = out0 * kDecim;                   // first PCM frame
  if (threadIdx.x < kTaps) taps[threadIdx.x] = tab->taps[threadIdx.x];

  const int16_t* __restrict__ pcm = tr.pcm;
  const uint32_t ch = tr.channels;
  const float chf = (float)ch;
  // frames past the end of the track contribute +0: acc + 0*h == acc, which is what the
  // reference's "start + j < n_samples" loop bound (resample.c:40) amounts to.
  if (ch == 1) {
    const int4* __restrict__ v = reinterpret_cast<const int4*>(pcm + in0);   // 16-byte aligned: in0 % 8 == 0
    for (int i = threadIdx.x; i < (kTileIn + 7) / 8; i += kK0Threads) {
      const uint64_t f = in0 + (uint64_t)i * 8;
      int4 q = make_int4(0, 0, 0, 0);
      if (f < tr.n_pcm) q = __ldg(v + i);      // buffers carry >= 16 bytes of slack past n_pcm
      const int w[4] = {q.x, q.y, q.z, q.w};
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        const int lo = (int)(int16_t)(w[k] & 0xffff), hi = w[k] >> 16;
        const int idx = i * 8 + 2 * k;
        if (idx < kTileIn) mono[pad32(idx)] = (f + 2 * k < tr.n_pcm) ? downmix_value(lo, chf) : 0.0f;
        if (idx + 1 < kTileIn) mono[pad32(idx + 1)] = (f + 2 * k + 1 < tr.n_pcm) ? downmix_value(hi, chf) : 0.0f;
      }
    }
  } else if (ch == 2) {
    const int4* __restrict__ v = reinterpret_cast<const int4*>(pcm + in0 * 2);
    for (int i = threadIdx.x; i < (kTileIn + 3) / 4; i += kK0Threads) {
      const uint64_t f = in0 + (uint64_t)i * 4;
      int4 q = make_int4(0, 0, 0, 0);
      if (f < tr.n_pcm) q = __ldg(v + i);
      const int w[4] = {q.x, q.y, q.z, q.w};
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        const int sum = (int)(int16_t)(w[k] & 0xffff) + (w[k] >> 16);
        const int idx = i * 4 + k;
        if (idx < kTileIn) mono[pad32(idx)] = (f + k < tr.n_pcm) ? downmix_value(sum, chf) : 0.0f;
      }
    }
  } else {
    for (int i = threadIdx.x; i < kTileIn; i += kK0Threads) {
      const uint64_t f = in0 + i;
      float m = 0.0f;
      if (f < tr.n_pcm) {
        int sum = 0;
        for (uint32_t c = 0; c < ch; ++c) sum += pcm[f * ch + c];
        m = downmix_value(sum, chf);
      }
      mono[pad32(i)] = m;
    }
  }
  __syncthreads();

  float* __restrict__ out = s5512 + tr.s_off;
#pragma unroll
  for (int m = 0; m < kK0OutPerThread; ++m) {
    const int o = threadIdx.x + m * kK0Threads;
    const uint64_t oi = out0 + o;
    if (oi >= tr.n5512) continue;
    float acc = 0.0f;   // resample.c:39-43: res += in[start + j] * h[j], j ascending
#pragma unroll
    for (int j = 0; j < kTaps; ++j) acc = __fadd_rn(acc, __fmul_rn(mono[pad32(o * kDecim + j)], taps[j]));
    out[oi] = acc;
  }
}

void launch_k0_resample(const TrackDev* tracks, const uint32_t* tile_prefix, uint32_t n_tracks, uint32_t n_tiles,
                        const Tables* tab, float* s5512, cudaStream_t st) {
  if (n_tiles == 0) return;
  k0_resample_kernel<<<n_tiles, kK0Threads, 0, st>>>(tracks, tile_prefix, n_tracks, tab, s5512);
}

}  // namespace mnemo
