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
    // s_off and chunk starts are multiples of 4 floats: 128-bit loads, 8 in flight per thread
    const float4* __restrict__ p4 = reinterpret_cast<const float4*>(p);
    const uint32_t n4 = cnt >> 2;
    uint32_t i = 0;
    for (; i + 8 <= n4; i += 8) {
      float4 q[8];
#pragma unroll
      for (int u = 0; u < 8; ++u) q[u] = __ldg(p4 + i + u);
#pragma unroll
      for (int u = 0; u < 8; ++u) {
        const float x[4] = {__fmul_rn(q[u].x, q[u].x), __fmul_rn(q[u].y, q[u].y), __fmul_rn(q[u].z, q[u].z),
                            __fmul_rn(q[u].w, q[u].w)};
#pragma unroll
        for (int k = 0; k < 4; ++k)
#pragma unroll
          for (int j = 0; j < kSumCands; ++j)
#pragma unroll
            for (int par = 0; par < 2; ++par) acc[j][par] = __fadd_rn(acc[j][par], x[k]);
      }
    }
    for (; i < n4; ++i) {
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

// D: one warp per track.  Within one binade the effect of a chunk is "m += delta[parity(m)]", and
// such maps compose associatively, so 32 chunks are stitched at once with a warp scan; the first
// chunk that would leave the binade (or whose candidates missed) is replayed sequentially from
// shared memory and the walk resumes behind it.
__device__ __forceinline__ void compose(uint32_t& a0, uint32_t& a1, uint32_t b0, uint32_t b1) {
  // (a after b)(p) = b[p] + a[(p + b[p]) & 1]
  const uint32_t n0 = b0 + ((b0 & 1u) ? a1 : a0);
  const uint32_t n1 = b1 + ((b1 & 1u) ? a0 : a1);
  a0 = n0;
  a1 = n1;
}

__global__ void __launch_bounds__(32) sumsq_stitch_kernel(const TrackDev* __restrict__ tracks,
                                                          const float* __restrict__ s5512,
                                                          const SumChunkRec* __restrict__ recs,
                                                          float* __restrict__ sumsq, float* __restrict__ rms_out) {
  __shared__ __align__(16) float sq[kSumChunk];
  const TrackDev tr = tracks[blockIdx.x];
  const int lane = threadIdx.x;
  const float* __restrict__ s = s5512 + tr.s_off;
  uint32_t S = 0;   // bits of the float accumulator (non-negative, so bit order == value order)
  for (uint32_t base = 0; base < tr.n_chunks; base += 32) {
    const uint32_t lim = min(32u, tr.n_chunks - base);
    SumChunkRec rec;
    rec.e_hi = 0;
    if ((uint32_t)lane < lim) rec = recs[tr.chunk_off + base + lane];
    uint32_t done = 0;
    while (done < lim) {
      const uint32_t e = S >> 23;
      const uint32_t m = (S & 0x7fffffu) | 0x800000u;
      const uint32_t j = rec.e_hi - e;   // candidate index (unsigned wrap = no match)
      uint32_t d0 = kSumInvalid, d1 = kSumInvalid;
#pragma unroll
      for (int c = 0; c < kSumCands; ++c)
        if (j == (uint32_t)c) {
          d0 = rec.delta[c][0];
          d1 = rec.delta[c][1];
        }
      const bool mine = (uint32_t)lane >= done && (uint32_t)lane < lim;
      const bool ok = mine && e >= 1 && e < 255 && rec.e_hi && d0 != kSumInvalid && d1 != kSumInvalid;
      uint32_t a0 = ok ? d0 : 0u, a1 = ok ? d1 : 0u;
#pragma unroll
      for (int o = 1; o < 32; o <<= 1) {
        const uint32_t b0 = __shfl_up_sync(0xffffffffu, a0, o), b1 = __shfl_up_sync(0xffffffffu, a1, o);
        if (lane >= o) compose(a0, a1, b0, b1);
      }
      const uint32_t tot = (m & 1u) ? a1 : a0;   // increment through this lane's chunk
      const uint32_t bad = __ballot_sync(0xffffffffu, mine && (!ok || m + tot >= 0x1000000u));
      const uint32_t first_bad = bad ? (uint32_t)(__ffs(bad) - 1) : lim;
      if (first_bad > done) S += __shfl_sync(0xffffffffu, tot, first_bad - 1);
      done = first_bad;
      if (done < lim) {   // sequential replay of chunk base+done
        const uint32_t first = (base + done) * kSumChunk;
        const uint32_t cnt = min((uint32_t)kSumChunk, tr.n5512 - first);
        for (uint32_t i = lane; i < (uint32_t)kSumChunk; i += 32) {
          const float v = i < cnt ? s[first + i] : 0.0f;
          sq[i] = __fmul_rn(v, v);   // past the end: + 0.0f leaves the sum unchanged
        }
        __syncwarp();
        float acc = __uint_as_float(S);
        const float4* q4 = reinterpret_cast<const float4*>(sq);
        const uint32_t n4 = (cnt + 3) >> 2;
#pragma unroll 4
        for (uint32_t i = 0; i < n4; ++i) {
          const float4 q = q4[i];
          acc = __fadd_rn(__fadd_rn(__fadd_rn(__fadd_rn(acc, q.x), q.y), q.z), q.w);
        }
        S = __float_as_uint(acc);
        __syncwarp();
        ++done;
      }
    }
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
