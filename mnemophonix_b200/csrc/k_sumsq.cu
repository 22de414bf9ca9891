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
//   C  per chunk (one warp): each lane simulates 128 samples with real f32 adds from the bottom of
//      each of the 3 binades the estimate allows x 2 parities; these maps "m += delta[parity]"
//      compose associatively, so a warp scan yields the chunk's 6 candidate increments
//   C' per group of 32 chunks (one warp): the 32 chunk maps composed once more, for the 3 binades the group's
//      estimates allow -> 6 candidate increments per group
//   D  per track, one warp: walks the groups; a group whose map exists for the accumulator's current binade
//      and keeps it inside that binade is taken in one step.  Otherwise (about once per binade the sum climbs
//      through) the group is stitched chunk by chunk with a warp scan, and the chunk that crosses the binade is
//      replayed slice-wise from the per-lane maps step C kept, the crossing 128-sample slice sequentially.
//      Then rms = (float)((double)sqrtf(S/(float)n) * 10.0) clamped.
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

// B: one CTA per track, exclusive scan over the track's chunks; also the inclusive value (prefix at the chunk's
// end), which is what bounds the binades the accumulator can visit inside the chunk.
constexpr int kPrefixThreads = 1024;
__global__ void __launch_bounds__(kPrefixThreads) sumsq_prefix_kernel(const TrackDev* __restrict__ tracks,
                                                                      const double* __restrict__ chunk_true,
                                                                      double* __restrict__ chunk_prefix,
                                                                      double* __restrict__ chunk_prefix_end) {
  __shared__ double warp_tot[32];
  __shared__ double s_carry;
  const TrackDev tr = tracks[blockIdx.x];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  if (threadIdx.x == 0) s_carry = 0.0;
  __syncthreads();
  for (uint32_t base = 0; base < tr.n_chunks; base += kPrefixThreads) {
    const uint32_t i = base + threadIdx.x;
    const double v = i < tr.n_chunks ? chunk_true[tr.chunk_off + i] : 0.0;
    double incl = v;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const double y = __shfl_up_sync(0xffffffffu, incl, o);
      if (lane >= o) incl += y;
    }
    if (lane == 31) warp_tot[warp] = incl;
    __syncthreads();
    double before = s_carry;
    for (int w = 0; w < warp; ++w) before += warp_tot[w];   // fixed order: the estimate is deterministic
    if (i < tr.n_chunks) {
      chunk_prefix[tr.chunk_off + i] = before + (incl - v);
      chunk_prefix_end[tr.chunk_off + i] = before + incl;
    }
    __syncthreads();
    if (threadIdx.x == kPrefixThreads - 1) s_carry = before + incl;
    __syncthreads();
  }
}

// ---- building blocks shared by C and D ----
constexpr int kSubChunk = kSumChunk / 32;   // 128 samples per lane

// (a after b)(p) = b[p] + a[(p + b[p]) & 1]: composition of two "m += delta[parity(m)]" maps
__device__ __forceinline__ void compose(uint32_t& a0, uint32_t& a1, uint32_t b0, uint32_t b1) {
  const uint32_t n0 = b0 + ((b0 & 1u) ? a1 : a0);
  const uint32_t n1 = b1 + ((b1 & 1u) ? a0 : a1);
  a0 = n0;
  a1 = n1;
}

// Inclusive scan of the per-lane maps over the warp (lane order = sample order).
__device__ __forceinline__ void scan_maps(uint32_t& a0, uint32_t& a1, int lane) {
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    const uint32_t b0 = __shfl_up_sync(0xffffffffu, a0, o), b1 = __shfl_up_sync(0xffffffffu, a1, o);
    if (lane >= o) compose(a0, a1, b0, b1);
  }
}

// One lane simulates its n (<= 128, multiple of 4 except at the track end) samples with real f32 adds
// from the bottom of NC candidate binades x 2 parities.  d[j][par] = mantissa increment, or
// kSumInvalid if that accumulator left its binade.
template <int NC>
__device__ __forceinline__ void lane_sim(const float* __restrict__ p, uint32_t n, const uint32_t (&e)[NC],
                                         uint32_t (&d)[NC][2]) {
  float acc[NC][2];
#pragma unroll
  for (int j = 0; j < NC; ++j)
#pragma unroll
    for (int par = 0; par < 2; ++par) acc[j][par] = __uint_as_float((e[j] << 23) | (uint32_t)par);
  const float4* __restrict__ p4 = reinterpret_cast<const float4*>(p);
  const uint32_t n4 = n >> 2;
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
        for (int j = 0; j < NC; ++j)
#pragma unroll
          for (int par = 0; par < 2; ++par) acc[j][par] = __fadd_rn(acc[j][par], x[k]);
    }
  }
  for (uint32_t k = i * 4; k < n; ++k) {
    const float x = __fmul_rn(p[k], p[k]);
#pragma unroll
    for (int j = 0; j < NC; ++j)
#pragma unroll
      for (int par = 0; par < 2; ++par) acc[j][par] = __fadd_rn(acc[j][par], x);
  }
#pragma unroll
  for (int j = 0; j < NC; ++j)
#pragma unroll
    for (int par = 0; par < 2; ++par) {
      const uint32_t start = (e[j] << 23) | (uint32_t)par, end = __float_as_uint(acc[j][par]);
      d[j][par] = (end >> 23) == e[j] ? end - start : kSumInvalid;
    }
}

// C: one WARP per chunk; each lane simulates 128 samples, the 32 maps are composed by a warp scan.
__global__ void __launch_bounds__(256) sumsq_sim_kernel(const TrackDev* __restrict__ tracks,
                                                        const uint32_t* __restrict__ chunk_track,
                                                        uint64_t n_chunks, const float* __restrict__ s5512,
                                                        const double* __restrict__ chunk_prefix_end,
                                                        SumChunkRec* __restrict__ recs,
                                                        uint32_t* __restrict__ slice_maps) {
  const uint64_t c = (uint64_t)blockIdx.x * 8 + (threadIdx.x >> 5);
  if (c >= n_chunks) return;
  const int lane = threadIdx.x & 31;
  const TrackDev tr = tracks[chunk_track[c]];
  const uint32_t first = (uint32_t)(c - tr.chunk_off) * kSumChunk;
  const uint32_t cnt = min((uint32_t)kSumChunk, tr.n5512 - first);
  const uint32_t lo = min(cnt, (uint32_t)lane * kSubChunk), n = min(cnt - lo, (uint32_t)kSubChunk);
  // The sequential float sum never exceeds the true sum by more than rounding noise, and in
  // practice stays above a quarter of it: candidates are the binade of the estimate at the END of the chunk
  // (so that a chunk which lifts the sum into the next binade still has a map for it) and the two below.
  const float est = __double2float_ru(chunk_prefix_end[c] * 1.000001);
  uint32_t e_hi = __float_as_uint(est) >> 23;
  if (!(est > 0.0f) || e_hi < 3 || e_hi > 254) e_hi = 0;
  SumChunkRec rec;
  rec.pad = 0;
  rec.e_hi = e_hi;
  if (e_hi) {
    const uint32_t e[kSumCands] = {e_hi, e_hi - 1, e_hi - 2};
    uint32_t d[kSumCands][2];
    lane_sim<kSumCands>(s5512 + tr.s_off + first + lo, n, e, d);
    // the per-lane (128-sample slice) maps are kept for the replay of chunks that cross a binade
    uint32_t* __restrict__ sm = slice_maps + c * (2 * kSumCands * 32);
#pragma unroll
    for (int j = 0; j < kSumCands; ++j) {
      sm[(2 * j) * 32 + lane] = d[j][0];
      sm[(2 * j + 1) * 32 + lane] = d[j][1];
    }
#pragma unroll
    for (int j = 0; j < kSumCands; ++j) {
      const bool bad = __any_sync(0xffffffffu, d[j][0] == kSumInvalid || d[j][1] == kSumInvalid);
      uint32_t a0 = d[j][0], a1 = d[j][1];
      scan_maps(a0, a1, lane);
      a0 = __shfl_sync(0xffffffffu, a0, 31);
      a1 = __shfl_sync(0xffffffffu, a1, 31);
      // valid only if the accumulator, started at the bottom of the binade, never leaves it
      rec.delta[j][0] = (!bad && a0 < 0x800000u) ? a0 : kSumInvalid;
      rec.delta[j][1] = (!bad && a1 + 1u < 0x800000u) ? a1 : kSumInvalid;
    }
  } else {
#pragma unroll
    for (int j = 0; j < kSumCands; ++j) rec.delta[j][0] = rec.delta[j][1] = kSumInvalid;
  }
  if (lane == 0) recs[c] = rec;
}

// C': one warp per group of 32 consecutive chunks of a track (launched one warp per chunk; only the warp of a
// group's first chunk works).  grecs[first chunk of the group] = the composed maps for the binades
// e_top, e_top-1, e_top-2, e_top = the highest estimate in the group.
__global__ void __launch_bounds__(256) sumsq_group_kernel(const TrackDev* __restrict__ tracks,
                                                          const uint32_t* __restrict__ chunk_track, uint64_t n_chunks,
                                                          const SumChunkRec* __restrict__ recs,
                                                          SumChunkRec* __restrict__ grecs) {
  const uint64_t c = (uint64_t)blockIdx.x * 8 + (threadIdx.x >> 5);
  if (c >= n_chunks) return;
  const int lane = threadIdx.x & 31;
  const TrackDev tr = tracks[chunk_track[c]];
  const uint32_t ci = (uint32_t)(c - tr.chunk_off);
  if (ci & 31u) return;
  const uint32_t lim = min(32u, tr.n_chunks - ci);
  SumChunkRec rec;
  rec.e_hi = 0;
  if ((uint32_t)lane < lim) rec = recs[c + lane];
  const uint32_t e_top = __reduce_max_sync(0xffffffffu, rec.e_hi);
  const bool all_est = !__any_sync(0xffffffffu, (uint32_t)lane < lim && rec.e_hi == 0);
  SumChunkRec g;
  g.pad = 0;
  g.e_hi = (all_est && e_top >= 3) ? e_top : 0;
#pragma unroll
  for (int k = 0; k < kSumCands; ++k) {
    const uint32_t j = rec.e_hi - (e_top - k);   // this chunk's candidate index for binade e_top - k
    uint32_t d0 = kSumInvalid, d1 = kSumInvalid;
#pragma unroll
    for (int q = 0; q < kSumCands; ++q)
      if (j == (uint32_t)q) {
        d0 = rec.delta[q][0];
        d1 = rec.delta[q][1];
      }
    const bool mine = (uint32_t)lane < lim;
    const bool ok = !mine || (d0 != kSumInvalid && d1 != kSumInvalid);
    const bool all_ok = __all_sync(0xffffffffu, ok);
    uint32_t a0 = mine ? d0 : 0u, a1 = mine ? d1 : 0u;   // identity map past the end of the track
    if (!all_ok) a0 = a1 = 0u;
    scan_maps(a0, a1, lane);
    a0 = __shfl_sync(0xffffffffu, a0, 31);
    a1 = __shfl_sync(0xffffffffu, a1, 31);
    // an increment of 2^23 or more cannot be applied inside one binade anyway
    g.delta[k][0] = (all_ok && g.e_hi && a0 < 0x800000u) ? a0 : kSumInvalid;
    g.delta[k][1] = (all_ok && g.e_hi && a1 < 0x800000u) ? a1 : kSumInvalid;
  }
  if (lane == 0) grecs[c] = g;
}

// D: one warp per track.  32 chunk maps are stitched at once with a warp scan; the first chunk that
// would leave the binade (or whose candidates missed) is replayed: lanes simulate their 128-sample
// slices under the CURRENT binade, a scan accepts the slices before the crossing, the crossing slice
// is added sequentially, and the loop continues in the new binade.
// sm / e_hi: the chunk's per-lane maps from step C (candidate j = binade e_hi - j), nullptr if none.
__device__ __forceinline__ uint32_t replay_chunk(uint32_t S, const float* __restrict__ p, uint32_t cnt, int lane,
                                                 const uint32_t* __restrict__ sm, uint32_t e_hi) {
  const uint32_t lo = min(cnt, (uint32_t)lane * kSubChunk), n = min(cnt - lo, (uint32_t)kSubChunk);
  const uint32_t n_slices = (cnt + kSubChunk - 1) / kSubChunk;
  uint32_t done = 0;
  bool may_sim = true;   // one fresh simulation per call at most: after it, slices without a stored map for the
                         // current binade are simply added sequentially (a start-of-track chunk climbs through
                         // twenty binades in its first few slices; re-simulating after each is far slower)
  while (done < n_slices) {
    const uint32_t e = S >> 23;
    uint32_t first_bad = done;   // without a usable binade: go straight to the sequential slice
    const uint32_t j = e_hi - e;
    const bool have_map = sm && e_hi && j < (uint32_t)kSumCands;   // warp-uniform
    if (e >= 1 && e < 255 && (have_map || may_sim)) {
      uint32_t d[1][2];
      if (have_map) {
        d[0][0] = sm[(2 * j) * 32 + lane];
        d[0][1] = sm[(2 * j + 1) * 32 + lane];
      } else {
        const uint32_t ee[1] = {e};
        lane_sim<1>(p + lo, n, ee, d);
        may_sim = false;
      }
      const bool mine = (uint32_t)lane >= done && (uint32_t)lane < n_slices;
      const bool ok = mine && d[0][0] != kSumInvalid && d[0][1] != kSumInvalid;
      uint32_t a0 = ok ? d[0][0] : 0u, a1 = ok ? d[0][1] : 0u;
      scan_maps(a0, a1, lane);
      const uint32_t m = (S & 0x7fffffu) | 0x800000u;
      const uint32_t tot = (m & 1u) ? a1 : a0;
      const uint32_t bad = __ballot_sync(0xffffffffu, mine && (!ok || m + tot >= 0x1000000u));
      first_bad = bad ? (uint32_t)(__ffs(bad) - 1) : n_slices;
      if (first_bad > done) S += __shfl_sync(0xffffffffu, tot, first_bad - 1);
    }
    done = first_bad;
    if (done < n_slices) {   // the crossing slice, sequentially (every lane computes the same chain)
      const uint32_t s0 = done * kSubChunk, sn = min((uint32_t)kSubChunk, cnt - s0);
      float acc = __uint_as_float(S);
      const float4* __restrict__ q4 = reinterpret_cast<const float4*>(p + s0);
      uint32_t i = 0;
      for (; i + 8 <= (sn >> 2); i += 8) {
        float4 q[8];
#pragma unroll
        for (int u = 0; u < 8; ++u) q[u] = __ldg(q4 + i + u);
#pragma unroll
        for (int u = 0; u < 8; ++u) {
          acc = __fadd_rn(acc, __fmul_rn(q[u].x, q[u].x));
          acc = __fadd_rn(acc, __fmul_rn(q[u].y, q[u].y));
          acc = __fadd_rn(acc, __fmul_rn(q[u].z, q[u].z));
          acc = __fadd_rn(acc, __fmul_rn(q[u].w, q[u].w));
        }
      }
      for (uint32_t k = i * 4; k < sn; ++k) acc = __fadd_rn(acc, __fmul_rn(p[s0 + k], p[s0 + k]));
      S = __float_as_uint(acc);
      ++done;
    }
  }
  return S;
}

__global__ void __launch_bounds__(32) sumsq_stitch_kernel(const TrackDev* __restrict__ tracks,
                                                          const float* __restrict__ s5512,
                                                          const SumChunkRec* __restrict__ recs,
                                                          const SumChunkRec* __restrict__ grecs,
                                                          const uint32_t* __restrict__ slice_maps,
                                                          float* __restrict__ sumsq, float* __restrict__ rms_out) {
  const TrackDev tr = tracks[blockIdx.x];
  const int lane = threadIdx.x;
  const float* __restrict__ s = s5512 + tr.s_off;
  uint32_t S = 0;   // bits of the float accumulator (non-negative, so bit order == value order)
  const uint32_t n_groups = (tr.n_chunks + 31u) >> 5;
  SumChunkRec gnext;
  gnext.e_hi = 0;
  if ((uint32_t)lane < n_groups) gnext = grecs[tr.chunk_off + 32ull * lane];
  for (uint32_t gbase = 0; gbase < n_groups; gbase += 32) {
    const SumChunkRec grec = gnext;   // lane l holds the record of group gbase + l; the next 32 are on their way
    gnext.e_hi = 0;
    if (gbase + 32 + lane < n_groups) gnext = grecs[tr.chunk_off + 32ull * (gbase + 32 + lane)];
    const uint32_t glim = min(32u, n_groups - gbase);
    for (uint32_t gl = 0; gl < glim; ++gl) {
      // fast path: the whole group in one step
      const uint32_t e = S >> 23;
      const uint32_t m = (S & 0x7fffffu) | 0x800000u;
      const uint32_t ge = __shfl_sync(0xffffffffu, grec.e_hi, gl);
      const uint32_t k = ge - e;
      uint32_t pick = kSumInvalid;   // my lane's view of delta[k][parity of m], then broadcast from lane gl
#pragma unroll
      for (int q = 0; q < kSumCands; ++q)
        if (k == (uint32_t)q) pick = (m & 1u) ? grec.delta[q][1] : grec.delta[q][0];
      const uint32_t tot_g = __shfl_sync(0xffffffffu, pick, gl);
      if (ge && e >= 1 && e < 255 && k < (uint32_t)kSumCands && tot_g != kSumInvalid && m + tot_g < 0x1000000u) {
        S += tot_g;
        continue;
      }
      // slow path: chunk by chunk
      const uint32_t base = (gbase + gl) << 5;
      const uint32_t lim = min(32u, tr.n_chunks - base);
      SumChunkRec rec;
      rec.e_hi = 0;
      if ((uint32_t)lane < lim) rec = recs[tr.chunk_off + base + lane];
      uint32_t done = 0;
      while (done < lim) {
        const uint32_t e2 = S >> 23;
        const uint32_t m2 = (S & 0x7fffffu) | 0x800000u;
        const uint32_t j = rec.e_hi - e2;   // candidate index (unsigned wrap = no match)
        uint32_t d0 = kSumInvalid, d1 = kSumInvalid;
#pragma unroll
        for (int c = 0; c < kSumCands; ++c)
          if (j == (uint32_t)c) {
            d0 = rec.delta[c][0];
            d1 = rec.delta[c][1];
          }
        const bool mine = (uint32_t)lane >= done && (uint32_t)lane < lim;
        const bool ok = mine && e2 >= 1 && e2 < 255 && rec.e_hi && d0 != kSumInvalid && d1 != kSumInvalid;
        uint32_t a0 = ok ? d0 : 0u, a1 = ok ? d1 : 0u;
        scan_maps(a0, a1, lane);
        const uint32_t tot = (m2 & 1u) ? a1 : a0;   // increment through this lane's chunk
        const uint32_t bad = __ballot_sync(0xffffffffu, mine && (!ok || m2 + tot >= 0x1000000u));
        const uint32_t first_bad = bad ? (uint32_t)(__ffs(bad) - 1) : lim;
        if (first_bad > done) S += __shfl_sync(0xffffffffu, tot, first_bad - 1);
        done = first_bad;
        if (done < lim) {
          const uint32_t first = (base + done) * kSumChunk;
          const uint32_t e_hi_c = __shfl_sync(0xffffffffu, rec.e_hi, done);
          S = replay_chunk(S, s + first, min((uint32_t)kSumChunk, tr.n5512 - first), lane,
                           slice_maps + (tr.chunk_off + base + done) * (2ull * kSumCands * 32), e_hi_c);
          ++done;
        }
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
                  const float* s5512, double* chunk_true, double* chunk_prefix, double* chunk_prefix_end,
                  SumChunkRec* recs, SumChunkRec* grecs, uint32_t* slice_maps, float* sumsq, float* rms, cudaStream_t st,
                  int* n_launches) {
  if (n_tracks == 0) return;
  if (n_chunks) {
    const unsigned grid = (unsigned)((n_chunks + 7) / 8);
    sumsq_true_kernel<<<grid, 256, 0, st>>>(tracks, chunk_track, n_chunks, s5512, chunk_true);
    sumsq_prefix_kernel<<<n_tracks, kPrefixThreads, 0, st>>>(tracks, chunk_true, chunk_prefix, chunk_prefix_end);
    sumsq_sim_kernel<<<grid, 256, 0, st>>>(tracks, chunk_track, n_chunks, s5512, chunk_prefix_end, recs, slice_maps);
    sumsq_group_kernel<<<grid, 256, 0, st>>>(tracks, chunk_track, n_chunks, recs, grecs);
    *n_launches += 4;
  }
  sumsq_stitch_kernel<<<n_tracks, 32, 0, st>>>(tracks, s5512, recs, grecs, slice_maps, sumsq, rms);
  *n_launches += 1;
}

}  // namespace mnemo
