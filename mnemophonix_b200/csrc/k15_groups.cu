// k15_groups.cu — the per-image work that consecutive spectral images share, done once.
//
// Image j is the 128-frame window of band energies starting at frame 8j (spectralimages.c:116-123); it is
// scaled by its own maximum M_j (spectralimages.c:52-77) and Haar-transformed along time (haar.c:56-67).
// Consecutive images overlap by 120 frames, and the first three Haar levels only ever combine frames inside an
// aligned group of 8 (levels 1, 2, 3 pair frames (2i, 2i+1), then pairs of pairs, then the two quads of a
// group).  So for a group of 8 frames and a given maximum M, the 8 scaled pixels of every band and their three
// Haar levels — 4 + 2 + 1 detail coefficients and 1 low — are the same in each of the up to 16 images that
// contain the group and have that maximum.  The window maximum changes every few images only (4 to 6 distinct
// values among the 16 images of a group on the synthetic corpus), so computing one 1 KB "group record" per
// (group, distinct maximum) replaces two thirds of what the fingerprint kernel used to recompute per image:
// 4096 double-precision scale + logf evaluations and 3584 Haar steps.
//
//   group_max_kernel    gmax[g]  = largest energy of the group's 8 x 32 values
//   group_record_kernel one CTA per group: the maxima M_j of the (up to) 16 images j = g-15..g that contain
//                       it, their distinct values (warp match), slot[j][g - j] = which record image j reads,
//                       imax[j] = M_j (written by the image's first group), then one warp per distinct M:
//                       lane = band, record = {hi1[0..3]; hi2[0..1], hi3, lo3} per band.
// Records live at fixed slots, rec[(group * 16 + d) * 256 .. + 256): memory is reserved for 16 per group
// (the size of all images), only the slots in use are ever written or read.
// Arithmetic: identical to the reference's, operation for operation (see scale_pixel / haar_pair).
#include "fingerprint_math.cuh"

namespace mnemo {

__global__ void __launch_bounds__(256) group_max_kernel(const TrackDev* __restrict__ tracks,
                                                        const uint32_t* __restrict__ group_prefix, uint32_t n_tracks,
                                                        uint32_t n_groups, const float* __restrict__ bins,
                                                        uint32_t* __restrict__ gmax) {
  const uint32_t gg = blockIdx.x * 8 + (threadIdx.x >> 5);   // global group index
  if (gg >= n_groups) return;
  const int lane = threadIdx.x & 31;
  const uint32_t t = find_segment<uint32_t>(group_prefix, n_tracks, gg);
  const TrackDev tr = tracks[t];
  const uint32_t g = gg - group_prefix[t];
  const float4* __restrict__ src = reinterpret_cast<const float4*>(bins + (tr.frame_off + 8ull * g) * kBands);
  const float4 a = __ldg(src + lane), b = __ldg(src + 32 + lane);   // 8 frames x 32 bands = 64 float4
  // energies are >= +0, so unsigned bit order is value order (a NaN sorts above everything and poisons the
  // window, as it does in the reference's running maximum... the images it touches are dropped either way)
  uint32_t m = max(max(__float_as_uint(a.x), __float_as_uint(a.y)), max(__float_as_uint(a.z), __float_as_uint(a.w)));
  m = max(m, max(max(__float_as_uint(b.x), __float_as_uint(b.y)), max(__float_as_uint(b.z), __float_as_uint(b.w))));
  m = __reduce_max_sync(0xffffffffu, m);
  if (lane == 0) gmax[gg] = m;
}

template <int LOGF_VARIANT>
__global__ void __launch_bounds__(128) group_record_kernel(const TrackDev* __restrict__ tracks,
                                                           const uint32_t* __restrict__ group_prefix, uint32_t n_tracks,
                                                           const Tables* __restrict__ tab, const float* __restrict__ bins,
                                                           const uint32_t* __restrict__ gmax, float* __restrict__ imax,
                                                           uint8_t* __restrict__ slot, float* __restrict__ rec,
                                                           float* __restrict__ tap_px) {
  __shared__ __align__(128) double s_logtab[32];
  __shared__ uint32_t s_m[16];
  __shared__ uint32_t s_count;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const uint32_t gg = blockIdx.x;
  const uint32_t t = find_segment<uint32_t>(group_prefix, n_tracks, gg);
  const TrackDev tr = tracks[t];
  const uint32_t g = gg - group_prefix[t];
  if (threadIdx.x < 32) s_logtab[threadIdx.x] = c_logf_tab[threadIdx.x & 15][threadIdx.x >> 4];
  if (warp == 0) {
    // lane l < 16 <-> image j = g - 15 + l of this track
    const int j = (int)g - 15 + lane;
    const bool valid = lane < 16 && j >= 0 && (uint32_t)j < tr.n_images;
    uint32_t mb = 0;
    if (valid) {
      const uint32_t* __restrict__ gm = gmax + group_prefix[t] + j;   // image j = groups j .. j+15 (all complete)
#pragma unroll
      for (int k = 0; k < 16; ++k) mb = max(mb, __ldg(gm + k));
    }
    const uint32_t key = valid ? mb : (0x80000000u | (uint32_t)lane);   // energies never have the sign bit
    const uint32_t peers = __match_any_sync(0xffffffffu, key);
    const int leader = __ffs(peers) - 1;
    const bool is_first = valid && leader == lane;
    const uint32_t firsts = __ballot_sync(0xffffffffu, is_first);
    const uint32_t d = __popc(firsts & ((1u << leader) - 1u));
    if (valid) {
      slot[(tr.img_off + (uint32_t)j) * 16 + (15 - lane)] = (uint8_t)d;
      if (lane == 15) imax[tr.img_off + (uint32_t)j] = __uint_as_float(mb);
    }
    if (is_first) s_m[d] = mb;
    if (lane == 0) s_count = __popc(firsts);
  }
  __syncthreads();
  const uint32_t count = s_count;
  const uint32_t logtab_saddr = (uint32_t)__cvta_generic_to_shared(s_logtab);
  const float log256 = tab->log256, inv_log256 = tab->inv_log256;
  const float* __restrict__ src = bins + (tr.frame_off + 8ull * g) * kBands + lane;
  for (uint32_t d = warp; d < count; d += 4) {
    const float mx = __uint_as_float(s_m[d]);
    if (!(mx > 0.0f) || !(mx < __int_as_float(0x7f800000))) continue;   // such images are dropped, nobody reads this
    const double dmx = (double)mx;
    const double rmx = __drcp_rn(dmx);
    float v[8], p[8];
#pragma unroll
    for (int f = 0; f < 8; ++f) v[f] = __ldg(src + f * kBands);
    const uint64_t base = ((uint64_t)gg * 16 + d) * 256;
#pragma unroll
    for (int f = 0; f < 8; ++f) p[f] = scale_pixel<LOGF_VARIANT>(v[f], dmx, rmx, log256, inv_log256, logtab_saddr);
    if (tap_px) {
#pragma unroll
      for (int f = 0; f < 8; ++f) tap_px[base + f * kBands + lane] = p[f];
    }
    float4 h1, o;   // h1 = level-1 details; o = {level-2 details, level-3 detail, level-3 low}
    float a0, a1, a2, a3, b0, b1;
    haar_pair(p[0], p[1], a0, h1.x);
    haar_pair(p[2], p[3], a1, h1.y);
    haar_pair(p[4], p[5], a2, h1.z);
    haar_pair(p[6], p[7], a3, h1.w);
    haar_pair(a0, a1, b0, o.x);
    haar_pair(a2, a3, b1, o.y);
    haar_pair(b0, b1, o.w, o.z);
    float4* __restrict__ dst = reinterpret_cast<float4*>(rec + base + 8 * lane);
    dst[0] = h1;
    dst[1] = o;
  }
}

void launch_group_records(const TrackDev* tracks, const uint32_t* group_prefix, uint32_t n_tracks, uint32_t n_groups,
                          const Tables* tab, const float* bins, uint32_t* gmax, float* imax, uint8_t* slot, float* rec,
                          float* tap_px, int logf_variant, cudaStream_t st, int* n_launches) {
  if (n_groups == 0) return;
  group_max_kernel<<<(n_groups + 7) / 8, 256, 0, st>>>(tracks, group_prefix, n_tracks, n_groups, bins, gmax);
  if (logf_variant)
    group_record_kernel<1><<<n_groups, 128, 0, st>>>(tracks, group_prefix, n_tracks, tab, bins, gmax, imax, slot, rec,
                                                     tap_px);
  else
    group_record_kernel<0><<<n_groups, 128, 0, st>>>(tracks, group_prefix, n_tracks, tab, bins, gmax, imax, slot, rec,
                                                     tap_px);
  *n_launches += 2;
}

}  // namespace mnemo
