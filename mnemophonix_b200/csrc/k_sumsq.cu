// k_sumsq.cu — the normaliser's whole-track float32 SEQUENTIAL sum of squares, in parallel.
//
// Replaces normalize()'s first loop and rms computation (audionormalizer.c:6-20).  The
// reference adds fl(s*s) to a float accumulator one sample at a time; on a 2-hour track that
// sum is 9 % below the true value, so a tree reduction is not the same
// number.  It is reproduced exactly as follows.
//
// While the accumulator S stays inside one binade [2^e, 2^(e+1)), S = m * u with u = 2^(e-23)
// and adding x >= 0 is integer arithmetic on m: m += floor(x/u) + round-bit, where the
// round-bit depends on m only through its parity (ties-to-even).  So the effect of a chunk of
// samples on S is, for a given (binade, parity of the incoming m), a fixed integer increment.
//   A  per chunk: true sum of the squares in double           -> estimate of S at chunk starts
//   B  per track: exclusive prefix of those sums
//   C  per chunk: simulate the chunk with real f32 adds from the bottom of each of the 3
//      binades the estimate allows x 2 parities               -> 6 candidate increments
//   D  per track, one warp: walk the chunks; take the increment that matches the actual
//      (binade, parity); if the chunk crosses a binade (or the estimate missed), redo that one
//      chunk sequentially.  Then rms = (float)((double)sqrtf(S/(float)n) * 10.0) clamped.
// The result is bit-identical to the sequential loop by construction, for any input; the
// estimate only decides how often step D has to fall back to the sequential path.
#include "common.cuh"

namespace mnemo {

// A: one warp per chunk
__global__ void __launch_bounds__(256) sumsq_true_kernel(const TrackDev* __restrict__ tracks,
                                                         const uint32_t* __restrict__ chunk_track,
                                                         uint64_t n_chunks, const float* __restrict__ s5512,
                                                         double* __restrict__ chunk_true) {
  const uint64_t c = (uint64_t)blockIdx.x * 8 + (threadIdx.x >> 5);
  if (c >= n_chunks) return;
  const int lane = threadIdx.x & 31;
  const TrackDev tr = tracks[chunk_track[c]];
  const uint32_t first = (uint32_t)(c - tr.chunk_off) * kSumChunk;
  const uint32_t cnt = min((uint32_t)kSumChunk, tr.n5512 - first);
  const float* __restrict__ p = s5512 + tr.s_off + first;
  double acc = 0.0;
  for (uint32_t i = lane; i < cnt; i += 32) {
    const float v = p[i];
    acc += (double)__fmul_rn(v, v);
  }
#pragma unroll
  for (int o = 16; o; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
  if (lane == 0) chunk_true[c] = acc;
}

// B: one warp per track, exclusive scan over the track's chunks
__global__ void __launch_bounds__(32) sumsq_prefix_kernel(const TrackDev* __restrict__ tracks,
                                                          const double* __restrict__ chunk_true,
                                                          double* __restrict__ chunk_prefix) {
  const TrackDev tr = tracks[blockIdx.x];
  const int lane = threadIdx.x;
  double carry = 0.0;
  for (uint32_t base = 0; base < tr.n_chunks; base += 32) {
    const uint32_t i = base + lane;
    double v = i < tr.n_chunks ? chunk_true[tr.chunk_off + i] : 0.0;
    double incl = v;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      double y = __shfl_up_sync(0xffffffffu, incl, o);
      if (lane >= o) incl += y;
    }
    if (i < tr.n_chunks) chunk_prefix[tr.chunk_off + i] = carry + (incl - v);
    carry += __shfl_sync(0xffffffffu, incl, 31);
  }
}

// C: one thread per chunk, 6 independent float accumulators
__global__ void __launch_bounds__(128) sumsq_sim_kernel(const TrackDev* __restrict__ tracks,
                                                        const uint32_t* __restrict__ chunk_track,
                                                        uint64_t n_chunks, const float* __restrict__ s5512,
                                                        const double* __restrict__ chunk_prefix,
                                                        SumChunkRec* __restrict__ recs) {
  const uint64_t c = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= n_chunks) return;
  const TrackDev tr = tracks[chunk_track[c]];
  const uint32_t first = (uint32_t)(c - tr.chunk_off) * kSumChunk;
  const uint32_t cnt = min((uint32_t)kSumChunk, tr.n5512 - first);
  const float* __restrict__ p = s5512 + tr.s_off + first;

  SumChunkRec rec;
  rec.pad = 0;
  // The sequential float sum never exceeds the true sum by more than rounding noise, and in
  // practice stays above a quarter of it: candidates are the estimate's binade and the two below.
  const float est = __double2float_ru(chunk_prefix[c] * 1.000001);
  uint32_t e_hi = __float_as_uint(est) >> 23;
  if (!(est > 0.0f) || e_hi < 3 || e_hi > 254) e_hi = 0;
  rec.e_hi = e_hi;
  float acc[kSumCands][2];
#pragma unroll
  for (int j = 0; j < kSumCands; ++j)
#pragma unroll
    for (int par = 0; par < 2; ++par) acc[j][par] = __uint_as_float(((e_hi - j) << 23) | (uint32_t)par);
  if (e_hi) {
    // s_off and chunk starts are multiples of 4 floats: 128-bit loads
    const float4* __restrict__ p4 = reinterpret_cast<const float4*>(p);
    const uint32_t n4 = cnt >> 2;
    for (uint32_t i = 0; i < n4; ++i) {
      const float4 q = __ldg(p4 + i);
      const float x[4] = {__fmul_rn(q.x, q.x), __fmul_rn(q.y, q.y), __fmul_rn(q.z, q.z), __fmul_rn(q.w, q.w)};
#pragma unroll
      for (int k = 0; k < 4; ++k)
#pragma unroll
        for (int j = 0; j < kSumCands; ++j)
#pragma unroll
          for (int par = 0; par < 2; ++par) acc[j][par] = __fadd_rn(acc[j][par], x[k]);
    }
    for (uint32_t i = n4 * 4; i < cnt; ++i) {
      const float x = __fmul_rn(p[i], p[i]);
#pragma unroll
      for (int j = 0; j < kSumCands; ++j)
#pragma unroll
        for (int par = 0; par < 2; ++par) acc[j][par] = __fadd_rn(acc[j][par], x);
    }
  }
#pragma unroll
  for (int j = 0; j < kSumCands; ++j)
#pragma unroll
    for (int par = 0; par < 2; ++par) {
      const uint32_t start = ((e_hi - j) << 23) | (uint32_t)par;
      const uint32_t end = __float_as_uint(acc[j][par]);
      // valid only while the simulated accumulator never left its binade
      rec.delta[j][par] = (e_hi && (end >> 23) == (e_hi - j)) ? end - start : kSumInvalid;
    }
  recs[c] = rec;
}

// D: one warp per track
__global__ void __launch_bounds__(32) sumsq_stitch_kernel(const TrackDev* __restrict__ tracks,
                                                          const float* __restrict__ s5512,
                                                          const SumChunkRec* __restrict__ recs,
                                                          float* __restrict__ sumsq, float* __restrict__ rms_out) {
  __shared__ SumChunkRec srec[32];
  const TrackDev tr = tracks[blockIdx.x];
  const int lane = threadIdx.x;
  const float* __restrict__ s = s5512 + tr.s_off;
  uint32_t S = 0;   // bits of the float accumulator (non-negative, so bit order == value order)
  for (uint32_t base = 0; base < tr.n_chunks; base += 32) {
    if (base + lane < tr.n_chunks) srec[lane] = recs[tr.chunk_off + base + lane];
    __syncwarp();
    const uint32_t lim = min(32u, tr.n_chunks - base);
    for (uint32_t l = 0; l < lim; ++l) {
      const SumChunkRec& r = srec[l];
      const uint32_t e = S >> 23;
      const uint32_t j = r.e_hi - e;    // candidate index (unsigned wrap = no match)
      bool done = false;
      if (r.e_hi && j < (uint32_t)kSumCands) {
        const uint32_t m = (S & 0x7fffffu) | 0x800000u;
        const uint32_t d = r.delta[j][m & 1u];
        if (d != kSumInvalid && m + d < 0x1000000u) {
          S += d;
          done = true;
        }
      }
      if (!done) {   // sequential replay of this chunk (all lanes compute the same chain)
        const uint32_t first = (base + l) * kSumChunk;
        const uint32_t cnt = min((uint32_t)kSumChunk, tr.n5512 - first);
        float acc = __uint_as_float(S);
        for (uint32_t i0 = 0; i0 < cnt; i0 += 32) {
          const uint32_t i = i0 + lane;
          const float v = i < cnt ? s[first + i] : 0.0f;
          const float x = __fmul_rn(v, v);
          const uint32_t nn = min(32u, cnt - i0);
          for (uint32_t k = 0; k < nn; ++k) acc = __fadd_rn(acc, __shfl_sync(0xffffffffu, x, k));
        }
        S = __float_as_uint(acc);
      }
    }
    __syncwarp();
  }
  if (lane == 0) {
    const float ss = __uint_as_float(S);
    // audionormalizer.c:15-20
    float rms = __double2float_rn(__dmul_rn((double)__fsqrt_rn(__fdiv_rn(ss, __uint2float_rn(tr.n5512))), 10.0));
    if ((double)rms < 0.1) rms = 0.1f;
    else if ((double)rms > 3.0) rms = 3.0f;
    sumsq[blockIdx.x] = ss;
    rms_out[blockIdx.x] = rms;
  }
}

void launch_sumsq(const TrackDev* tracks, uint32_t n_tracks, uint64_t n_chunks, const uint32_t* chunk_track,
                  const float* s5512, double* chunk_true, double* chunk_prefix, SumChunkRec* recs, float* sumsq,
                  float* rms, cudaStream_t st, int* n_launches) {
  if (n_tracks == 0) return;
  if (n_chunks) {
    sumsq_true_kernel<<<(unsigned)((n_chunks + 7) / 8), 256, 0, st>>>(tracks, chunk_track, n_chunks, s5512, chunk_true);
    sumsq_prefix_kernel<<<n_tracks, 32, 0, st>>>(tracks, chunk_true, chunk_prefix);
    sumsq_sim_kernel<<<(unsigned)((n_chunks + 127) / 128), 128, 0, st>>>(tracks, chunk_track, n_chunks, s5512,
                                                                          chunk_prefix, recs);
    *n_launches += 3;
  }
  sumsq_stitch_kernel<<<n_tracks, 32, 0, st>>>(tracks, s5512, recs, sumsq, rms);
  *n_launches += 1;
}

}  // namespace mnemo
