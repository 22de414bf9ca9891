// api_index.cu — C ABI (include/mnemo_cuda.h): context, host tables, indexing batches.
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <string>
#include <vector>

#include "../../include/mnemo_cuda.h"
#include "api_internal.h"

using namespace mnemo;

static const uint16_t k_permutations[kSigLen * kPermLen] = {
#include "../data/permutations_table.inc"
};

// ------------------------------------------------------------------ host tables ----
// The reference computes these lazily on the host with libm; so do we, with the same
// expressions (C promotion rules included), so every table is bit-identical to the one the
// reference builds on the same machine.

static void host_taps(float h[kTaps]) {   // resample.c:12-35
  for (int x = -15; x <= 15; x++) {
    if (x == 0) {
      h[15] = 0.125f;
      continue;
    }
    const float xs = (float)(x * 0.125);
    const double px = M_PI * (double)xs;
    const float sc = (float)((double)sinf((float)px) / px);
    const float d = (float)x - 15;
    const double a1 = 2 * M_PI * (double)d / 30;
    const double a2 = 4 * M_PI * (double)d / 30;
    const float bw = (float)(0.42 - 0.5 * (double)cosf((float)a1) + 0.08 * (double)cosf((float)a2));
    h[x + 15] = (float)(0.125 * (double)sc * (double)bw);
  }
}

static void host_hann(float* w) {   // hannwindow.c:8-12
  for (unsigned i = 0; i < (unsigned)kFrame; i++) {
    const float c = cosf((float)(2 * M_PI * i / (kFrame - 1)));
    const float om = 1 - c;
    w[i] = (float)(0.5 * (double)om);
  }
}

static void host_twiddles(float2* tw) {   // fft.c:49-51 at l = 2048
  for (int k = 0; k < 1024; k++) {
    const float kth = (float)(-2.0 * k * M_PI / 2048);
    tw[k] = make_float2(cosf(kth), sinf(kth));
  }
}

static void host_edges(uint16_t* e) {   // logbins.c:20-55
  const float lo = log2f(318), hi = log2f(2000);
  const float step = (hi - lo) / kBands;
  float cur = lo;
  for (int i = 0; i <= kBands; i++) {
    const float f = powf(2, cur);
    cur += step;
    int idx = (int)roundf((float)(1024.0 * (double)f / 2756.0));
    if (idx < 1) idx = 1;
    if (idx > 1024) idx = 1024;
    e[i] = (uint16_t)idx;
  }
}

// glibc logf, host restatement of the two builds (same constants as common.cuh)
static const double h_logf_tab[16][2] = {
    {0x1.661ec79f8f3bep+0, -0x1.57bf7808caadep-2}, {0x1.571ed4aaf883dp+0, -0x1.2bef0a7c06ddbp-2},
    {0x1.49539f0f010bp+0, -0x1.01eae7f513a67p-2},  {0x1.3c995b0b80385p+0, -0x1.b31d8a68224e9p-3},
    {0x1.30d190c8864a5p+0, -0x1.6574f0ac07758p-3}, {0x1.25e227b0b8eap+0, -0x1.1aa2bc79c81p-3},
    {0x1.1bb4a4a1a343fp+0, -0x1.a4e76ce8c0e5ep-4}, {0x1.12358f08ae5bap+0, -0x1.1973c5a611cccp-4},
    {0x1.0953f419900a7p+0, -0x1.252f438e10c1ep-5}, {0x1p+0, 0x0p+0},
    {0x1.e608cfd9a47acp-1, 0x1.aa5aa5df25984p-5},  {0x1.ca4b31f026aap-1, 0x1.c5e53aa362eb4p-4},
    {0x1.b2036576afce6p-1, 0x1.526e57720db08p-3},  {0x1.9c2d163a1aa2dp-1, 0x1.bc2860d22477p-3},
    {0x1.886e6037841edp-1, 0x1.1058bc8a07ee1p-2},  {0x1.767dcf5534862p-1, 0x1.4043057b6ee09p-2}};

static float host_logf_mirror(float x, int variant) {
  const double ln2 = 0x1.62e42fefa39efp-1;
  const double a0 = -0x1.00ea348b88334p-2, a1 = 0x1.5575b0be00b6ap-2, a2 = -0x1.ffffef20a4123p-2;
  uint32_t ix;
  memcpy(&ix, &x, 4);
  if (ix == 0x3f800000u) return 0.0f;
  const uint32_t tmp = ix - 0x3f330000u;
  const int i = (tmp >> 19) & 15;
  const int k = (int32_t)tmp >> 23;
  const uint32_t iz = ix - (tmp & 0xff800000u);
  float zf;
  memcpy(&zf, &iz, 4);
  const volatile double z = (double)zf;   // volatile: keep the compiler from contracting the baseline path
  const double invc = h_logf_tab[i][0], logc = h_logf_tab[i][1];
  if (variant) {
    const double y0 = fma((double)k, ln2, logc);
    const double r = fma(z, invc, -1.0);
    double y = fma(r, a1, a2);
    const double r2 = r * r;
    const double hi = r + y0;
    y = fma(r2, a0, y);
    return (float)fma(r2, y, hi);
  }
  volatile double p = z * invc;
  const double r = p - 1.0;
  p = (double)k * ln2;
  const double y0 = logc + p;
  p = r * r;
  const double r2 = p;
  p = a1 * r;
  double y = p + a2;
  p = a0 * r2;
  y = p + y;
  p = y * r2;
  const double q = y0 + r;
  return (float)(p + q);
}

static int build_tables(Tables* t, uint32_t* probes, uint32_t* matched, std::string* err) {
  memset(t, 0, sizeof(*t));
  host_taps(t->taps);
  host_hann(t->hann);
  std::vector<float2> tw(1024);
  host_twiddles(tw.data());
  for (int m = 0; m < 32; m++) t->tw64[m] = tw[32 * m];
  for (int h = 0; h < 2; h++)
    for (int slot = 0; slot < 31; slot++) {
      int s, q;
      if (slot < 1) s = 7, q = 0;
      else if (slot < 3) s = 8, q = slot - 1;
      else if (slot < 7) s = 9, q = slot - 3;
      else if (slot < 15) s = 10, q = slot - 7;
      else s = 11, q = slot - 15;
      for (int lane = 0; lane < 32; lane++) {
        const int i = lane + 32 * h;
        const int k = 64 * q + i;
        t->tw_hi[h][slot][lane] = tw[k << (11 - s)];
      }
    }
  host_edges(t->edges);
  t->log256 = logf((float)256.0);
  t->inv_log256 = 1.0f / t->log256;
  // Preconditions the kernels rely on.
  if (!(t->tw64[0].x == 1.0f && t->tw64[0].y == 0.0f)) {
    *err = "host libm: cosf/sinf(-0) is not (1, -0)";
    return -3;
  }
  if (t->edges[0] < 64 || t->edges[kBands] > 768) {
    *err = "host libm: band edges fall outside FFT bins [64, 768)";
    return -3;
  }
  for (int i = 0; i < kBands; i++) {
    if (t->edges[i + 1] <= t->edges[i]) {
      *err = "host libm: band edges are not increasing";
      return -3;
    }
    if (t->edges[i + 1] - t->edges[i] > k1_max_band_width()) {
      *err = "host libm: a band is wider than the spectral kernel's unrolled sum";
      return -3;
    }
  }
  // Which build of glibc's logf does this host run?  glibc selects the FMA variant when the CPU
  // has FMA and AVX2 (sysdeps/x86_64/fpu/multiarch/ifunc-fma.h); confirm against the real logf.
  __builtin_cpu_init();
  int variant = (__builtin_cpu_supports("fma") && __builtin_cpu_supports("avx2")) ? 1 : 0;
  uint32_t n = 0, ok[2] = {0, 0};
  for (uint32_t bits = 0x3f800000u; bits <= 0x43800000u; bits += 1021u) {   // [1, 256]
    float x;
    memcpy(&x, &bits, 4);
    const float ref = logf(x);
    ok[0] += host_logf_mirror(x, 0) == ref;
    ok[1] += host_logf_mirror(x, 1) == ref;
    n++;
  }
  if (ok[1 - variant] > ok[variant]) variant = 1 - variant;
  t->logf_variant = variant;
  *probes = n;
  *matched = ok[variant];
  return 0;
}

// ------------------------------------------------------------------ context ----

extern "C" int mnemo_create_ex(int device, void* stream, uint32_t flags, mnemo_ctx** out) {
  if (!out) return -3;
  *out = nullptr;
  mnemo_ctx* ctx = new (std::nothrow) mnemo_ctx();
  if (!ctx) return -1;
  int n_dev = 0;
  cudaError_t e = cudaGetDeviceCount(&n_dev);
  if (e != cudaSuccess || device < 0 || device >= n_dev) {
    fprintf(stderr, "mnemo_create: no usable CUDA device %d (%s)\n", device, cudaGetErrorString(e));
    delete ctx;
    return -3;
  }
  ctx->device = device;
  cudaDeviceProp prop;
  e = cudaGetDeviceProperties(&prop, device);
  if (e != cudaSuccess || prop.major < 10) {
    fprintf(stderr, "mnemo_create: device %d is not sm_100-class (no CPU or legacy-GPU fallback)\n", device);
    delete ctx;
    return -3;
  }
  ctx->sm_count = prop.multiProcessorCount;
  if (cudaSetDevice(device) != cudaSuccess) {
    delete ctx;
    return -3;
  }
  if (stream) {
    ctx->stream = (cudaStream_t)stream;
    ctx->own_stream = false;
  } else {
    if (cudaStreamCreateWithFlags(&ctx->stream, cudaStreamNonBlocking) != cudaSuccess) {
      delete ctx;
      return -3;
    }
    ctx->own_stream = true;
  }
  std::string err;
  int rc = build_tables(&ctx->h_tables, &ctx->logf_probes, &ctx->logf_matched, &err);
  if (rc != 0) {
    fprintf(stderr, "mnemo_create: %s\n", err.c_str());
    mnemo_destroy(ctx);
    return rc;
  }
  // inverse of the permutation table (permutations.c:7-1808): inv[b][p] = j with perm[p][j] == b, else 255
  // (16-bit entries: the kernel takes minima with the two-lane 16-bit SIMD instruction)
  std::vector<uint16_t> inv((size_t)(8192 + 1) * kSigLen, 255);   // + one all-255 row used as padding
  for (int p = 0; p < kSigLen; p++)
    for (int j = kPermLen - 1; j >= 0; j--) inv[(size_t)k_permutations[p * kPermLen + j] * kSigLen + p] = (uint16_t)j;
  if (cudaMalloc(&ctx->d_tables, sizeof(Tables)) != cudaSuccess ||
      cudaMalloc(&ctx->d_inv_perms, inv.size() * sizeof(uint16_t)) != cudaSuccess) {
    mnemo_destroy(ctx);
    return -1;
  }
  cudaMemcpy(ctx->d_tables, &ctx->h_tables, sizeof(Tables), cudaMemcpyHostToDevice);
  cudaMemcpy(ctx->d_inv_perms, inv.data(), inv.size() * sizeof(uint16_t), cudaMemcpyHostToDevice);
  k1_upload_constants(ctx->h_tables);
  k0_upload_constants(ctx->h_tables);
  ctx->fast_spectra = (flags & MNEMO_FAST_SPECTRA) != 0;
  if (ctx->fast_spectra) {   // tolerance mode of K1: its twiddle tables
    std::vector<float2> tw(k1_fast_table_float2());
    float2 w32[16];
    k1_fast_tables(tw.data(), w32);
    if (cudaMalloc(&ctx->d_tw_fast, tw.size() * sizeof(float2)) != cudaSuccess) {
      mnemo_destroy(ctx);
      return -1;
    }
    cudaMemcpy(ctx->d_tw_fast, tw.data(), tw.size() * sizeof(float2), cudaMemcpyHostToDevice);
    k1_fast_upload_constants(w32);
  }
  e = cudaDeviceSynchronize();
  if (e != cudaSuccess) {
    fprintf(stderr, "mnemo_create: %s\n", cudaGetErrorString(e));
    mnemo_destroy(ctx);
    return -3;
  }
  *out = ctx;
  return 0;
}

extern "C" int mnemo_create(int device, void* stream, mnemo_ctx** out) {
  // MNEMO_FAST_SPECTRA=1 opts a whole process (e.g. the CLI) into the tolerance mode; the default is the exact replay
  const char* env = getenv("MNEMO_FAST_SPECTRA");
  return mnemo_create_ex(device, stream, (env && atoi(env) > 0) ? MNEMO_FAST_SPECTRA : 0u, out);
}

extern "C" int mnemo_spectra_mode(const mnemo_ctx* ctx) { return ctx && ctx->fast_spectra ? 1 : 0; }

static void live_release(mnemo_ctx* ctx, bool keep_context = false) {
  mnemo_ctx::Live& lv = ctx->live;
  if (lv.exec) cudaGraphExecDestroy(lv.exec);
  if (lv.batch) mnemo_batch_destroy(lv.batch);
  if (lv.pin_in) cudaFreeHost(lv.pin_in);
  if (lv.pin_out) cudaFreeHost(lv.pin_out);
  mnemo_ctx* own = lv.own;
  lv = mnemo_ctx::Live();
  if (keep_context) lv.own = own;
  else if (own) mnemo_destroy(own);
}

extern "C" void mnemo_destroy(mnemo_ctx* ctx) {
  if (!ctx) return;
  cudaSetDevice(ctx->device);
  live_release(ctx);
  if (ctx->arena) cudaFree(ctx->arena);
  if (ctx->d_tw_fast) cudaFree(ctx->d_tw_fast);
  if (ctx->d_tables) cudaFree(ctx->d_tables);
  if (ctx->d_inv_perms) cudaFree(ctx->d_inv_perms);
  if (ctx->own_stream && ctx->stream) cudaStreamDestroy(ctx->stream);
  delete ctx;
}

extern "C" const char* mnemo_last_error(const mnemo_ctx* ctx) { return ctx ? ctx->err.c_str() : "no context"; }

extern "C" int mnemo_logf_variant(const mnemo_ctx* ctx, uint32_t* probes, uint32_t* matched) {
  if (probes) *probes = ctx->logf_probes;
  if (matched) *matched = ctx->logf_matched;
  return ctx->h_tables.logf_variant;
}

int mnemo_fail(mnemo_ctx* ctx, cudaError_t e, const char* what) {
  char buf[256];
  snprintf(buf, sizeof buf, "%s: %s", what, cudaGetErrorString(e));
  ctx->err = buf;
  return e == cudaErrorMemoryAllocation ? -1 : -3;
}

// ------------------------------------------------------------------ sizes ----

extern "C" uint32_t mnemo_samples_5512(uint64_t n_pcm_frames) { return (uint32_t)(n_pcm_frames / kDecim); }
extern "C" uint32_t mnemo_frames(uint32_t n) {   // fingerprinting.c:42, spectralimages.c:37-39,128
  if (n < (uint32_t)kFrame) return 0;
  const uint32_t f = 1 + (n - kFrame) / kHop;
  return f < (uint32_t)kImgW ? 0 : f;
}
extern "C" uint32_t mnemo_images(uint32_t n) {   // spectralimages.c:47-49
  const uint32_t f = mnemo_frames(n);
  return f ? 1 + (f - kImgW) / kImgStep : 0;
}

// ------------------------------------------------------------------ batches ----

static uint64_t align_up(uint64_t v, uint64_t a) { return (v + a - 1) / a * a; }

template <typename T>
static cudaError_t dalloc(T** p, uint64_t count) {
  *p = nullptr;
  return cudaMalloc((void**)p, (count ? count : 1) * sizeof(T));
}

// hands out 256-byte aligned pieces of one allocation (base == nullptr: only adds up the sizes)
struct Carver {
  uint8_t* base;
  uint64_t off;
  template <typename T>
  void take(T** p, uint64_t count) {
    const uint64_t bytes = align_up((count ? count : 1) * sizeof(T), 256);
    if (base) *p = reinterpret_cast<T*>(base + off);
    off += bytes;
  }
};

static int batch_create_common(mnemo_ctx* ctx, const mnemo_track* tracks, uint32_t n_tracks, bool prenormalized,
                               mnemo_batch** out, bool share_arena = true) {
  if (!ctx || !out || (n_tracks && !tracks)) return -3;
  *out = nullptr;
  cudaSetDevice(ctx->device);
  mnemo_batch* b = new (std::nothrow) mnemo_batch();
  if (!b) return -1;
  b->ctx = ctx;
  b->n_tracks = n_tracks;
  b->prenormalized = prenormalized;
  b->h_tracks.resize(n_tracks);
  b->k0_tile_prefix.assign(n_tracks + 1, 0);
  b->k1_tile_prefix.assign(n_tracks + 1, 0);
  b->img_prefix.assign(n_tracks + 1, 0);
  b->pcm_off.assign(n_tracks + 1, 0);
  b->group_prefix.assign(n_tracks + 1, 0);
  const uint32_t k0_out = (uint32_t)k0_tile_outputs(), k1_fr = (uint32_t)k1_tile_frames();
  uint64_t s_off = 0, frame_off = 0, chunk_off = 0;
  for (uint32_t t = 0; t < n_tracks; t++) {
    TrackDev& d = b->h_tracks[t];
    memset(&d, 0, sizeof d);
    if (tracks[t].channels == 0 && !prenormalized) {
      delete b;
      ctx->err = "track with zero channels";
      return -4;
    }
    d.n_pcm = tracks[t].n_frames;
    d.channels = prenormalized ? 1 : tracks[t].channels;
    const uint64_t n5 = prenormalized ? tracks[t].n_frames : tracks[t].n_frames / kDecim;
    if (n5 > 0xffffffffull) {
      delete b;
      ctx->err = "track longer than 2^32 samples at 5512 Hz";
      return -3;
    }
    d.n5512 = (uint32_t)n5;
    d.n_frames = mnemo_frames(d.n5512);
    d.n_images = mnemo_images(d.n5512);
    d.prenormalized = prenormalized ? 1 : 0;
    d.n_chunks = (prenormalized || !d.n_frames) ? 0 : (d.n5512 + kSumChunk - 1) / kSumChunk;
    d.s_off = s_off;
    d.frame_off = frame_off;
    d.img_off = b->img_prefix[t];
    d.chunk_off = chunk_off;
    d.group_off = b->group_prefix[t];
    // images only ever touch complete groups: the last image ends at frame 8 * (n_images + 15) - 1 < n_frames
    b->group_prefix[t + 1] = b->group_prefix[t] + (d.n_images ? d.n_images + 15 : 0);
    s_off += align_up(d.n5512, 64);
    frame_off += d.n_frames;
    chunk_off += d.n_chunks;
    b->img_prefix[t + 1] = b->img_prefix[t] + d.n_images;
    // tracks that cannot be fingerprinted are skipped by every kernel
    b->k0_tile_prefix[t + 1] = b->k0_tile_prefix[t] + ((d.n_frames && !prenormalized) ? (d.n5512 + k0_out - 1) / k0_out : 0);
    b->k1_tile_prefix[t + 1] = b->k1_tile_prefix[t] + (d.n_frames + k1_fr - 1) / k1_fr;
    const uint64_t bytes = prenormalized ? 0 : tracks[t].n_frames * tracks[t].channels * 2ull;
    b->pcm_off[t + 1] = b->pcm_off[t] + align_up(bytes + 64, 256);
  }
  b->n_samples = s_off;
  b->n_frames = frame_off;
  b->n_images = b->img_prefix[n_tracks];
  b->n_chunks = chunk_off;
  b->n_groups = b->group_prefix[n_tracks];
  if (b->n_images >= 0xffffffffull || b->n_frames * (uint64_t)kBands >= (1ull << 40)) {
    delete b;
    ctx->err = "batch too large";
    return -3;
  }
  std::vector<uint32_t> chunk_track(b->n_chunks);
  for (uint32_t t = 0; t < n_tracks; t++)
    for (uint32_t c = 0; c < b->h_tracks[t].n_chunks; c++) chunk_track[b->h_tracks[t].chunk_off + c] = t;
  // tracks too small: no chunks needed either (rms unused) — keep them, they are cheap

  const uint32_t n_scan_blocks = (uint32_t)((b->n_images + 2047) / 2048);
  cudaError_t e = cudaSuccess;
#define TRY(x)                      \
  do {                              \
    if (e == cudaSuccess) e = (x);  \
  } while (0)
  // every buffer of the batch comes out of one arena: first pass sizes it, second pass hands out the pieces
  auto layout = [&](Carver& c) {
    c.take(&b->d_pcm, b->pcm_off[n_tracks]);
    c.take(&b->d_tracks, n_tracks);
    c.take(&b->d_k0_prefix, n_tracks + 1);
    c.take(&b->d_k1_prefix, n_tracks + 1);
    c.take(&b->d_img_prefix, n_tracks + 1);
    c.take(&b->d_chunk_track, b->n_chunks);
    c.take(&b->d_s5512, b->n_samples);
    if (!prenormalized) c.take(&b->d_norm, b->n_samples);
    c.take(&b->d_chunk_true, b->n_chunks);
    c.take(&b->d_chunk_prefix, b->n_chunks);
    c.take(&b->d_chunk_prefix_end, b->n_chunks);
    c.take(&b->d_recs, b->n_chunks);
    c.take(&b->d_grecs, b->n_chunks);                                  // used at the first chunk of every group of 32
    c.take(&b->d_slice_maps, b->n_chunks * (2ull * kSumCands * 32));   // 768 B per chunk
    c.take(&b->d_sumsq, n_tracks);
    c.take(&b->d_rms, n_tracks);
    c.take(&b->d_bins, b->n_frames * kBands);
    c.take(&b->d_group_prefix, n_tracks + 1);
    c.take(&b->d_gmax, b->n_groups);
    c.take(&b->d_imax, b->n_images);
    c.take(&b->d_slot, b->n_images * 16);
    c.take(&b->d_grec, b->n_groups * 16 * 256);   // room for 16 records per group; traffic follows the slots in use
    c.take(&b->d_sig_dense, b->n_images * kSigLen);
    c.take(&b->d_keep, b->n_images);
    c.take(&b->d_pos, b->n_images + 1);
    c.take(&b->d_block_sums, (uint64_t)n_scan_blocks + 1);
    c.take(&b->d_sig_out, b->n_images * kSigLen);
    c.take(&b->d_n_sigs, n_tracks);
  };
  Carver sizing{nullptr, 0};
  layout(sizing);
  const uint64_t arena_bytes = sizing.off;
  if (share_arena) {
    std::lock_guard<std::mutex> lock(ctx->arena_mu);
    if (!ctx->arena_busy) {
      if (ctx->arena_bytes < arena_bytes) {
        if (ctx->arena) cudaFree(ctx->arena);
        ctx->arena = nullptr;
        ctx->arena_bytes = 0;
        e = cudaMalloc((void**)&ctx->arena, arena_bytes);
        if (e == cudaSuccess) ctx->arena_bytes = arena_bytes;
      }
      if (e == cudaSuccess) {
        ctx->arena_busy = true;
        b->arena = ctx->arena;
        b->arena_of_ctx = true;
      }
    }
  }
  if (e == cudaSuccess && !b->arena) e = cudaMalloc((void**)&b->arena, arena_bytes);
  if (e == cudaSuccess) {
    Carver carve{b->arena, 0};
    layout(carve);
  }
  if (e != cudaSuccess) {
    int rc = mnemo_fail(ctx, e, "mnemo_batch_create");
    mnemo_batch_destroy(b);
    return rc;
  }
  if (prenormalized) b->d_norm = b->d_s5512;
  for (uint32_t t = 0; t < n_tracks; t++)
    b->h_tracks[t].pcm = prenormalized ? nullptr : reinterpret_cast<const int16_t*>(b->d_pcm + b->pcm_off[t]);
  cudaStream_t st = ctx->stream;
  TRY(cudaMemcpyAsync(b->d_tracks, b->h_tracks.data(), sizeof(TrackDev) * n_tracks, cudaMemcpyHostToDevice, st));
  TRY(cudaMemcpyAsync(b->d_k0_prefix, b->k0_tile_prefix.data(), 4ull * (n_tracks + 1), cudaMemcpyHostToDevice, st));
  TRY(cudaMemcpyAsync(b->d_k1_prefix, b->k1_tile_prefix.data(), 4ull * (n_tracks + 1), cudaMemcpyHostToDevice, st));
  TRY(cudaMemcpyAsync(b->d_img_prefix, b->img_prefix.data(), 8ull * (n_tracks + 1), cudaMemcpyHostToDevice, st));
  TRY(cudaMemcpyAsync(b->d_group_prefix, b->group_prefix.data(), 4ull * (n_tracks + 1), cudaMemcpyHostToDevice, st));
  if (b->n_chunks)
    TRY(cudaMemcpyAsync(b->d_chunk_track, chunk_track.data(), 4ull * b->n_chunks, cudaMemcpyHostToDevice, st));
  // the slack after each track's PCM is read (masked) by the 128-bit loads: keep it defined
  TRY(cudaMemsetAsync(b->d_pcm, 0, b->pcm_off[n_tracks] ? b->pcm_off[n_tracks] : 1, st));
  TRY(cudaMemsetAsync(b->d_keep, 0, b->n_images ? b->n_images : 1, st));
  TRY(cudaStreamSynchronize(st));
#undef TRY
  if (e != cudaSuccess) {
    int rc = mnemo_fail(ctx, e, "mnemo_batch_create");
    mnemo_batch_destroy(b);
    return rc;
  }
  *out = b;
  return 0;
}

extern "C" int mnemo_batch_create(mnemo_ctx* ctx, const mnemo_track* tracks, uint32_t n_tracks, mnemo_batch** out) {
  return batch_create_common(ctx, tracks, n_tracks, false, out);
}

extern "C" void mnemo_batch_destroy(mnemo_batch* b) {
  if (!b) return;
  cudaSetDevice(b->ctx->device);
  cudaStreamSynchronize(b->ctx->stream);
  if (b->arena_of_ctx) {
    std::lock_guard<std::mutex> lock(b->ctx->arena_mu);
    b->ctx->arena_busy = false;   // stays allocated for the next batch
  } else if (b->arena) {
    cudaFree(b->arena);
  }
  void* ptrs[] = {b->d_tap_px, b->d_tap_raw, b->d_tap_img, b->d_tap_haar};
  for (void* p : ptrs)
    if (p) cudaFree(p);
  for (int i = 0; i < 6; i++)
    if (b->ev[i]) cudaEventDestroy(b->ev[i]);
  for (cudaEvent_t ev : b->up_ev)
    if (ev) cudaEventDestroy(ev);
  delete b;
}

extern "C" int mnemo_batch_upload(mnemo_batch* b, const mnemo_track* tracks) {
  if (!b || !tracks) return -3;
  cudaSetDevice(b->ctx->device);
  for (uint32_t t = 0; t < b->n_tracks; t++) {
    if (b->prenormalized) {
      const uint64_t bytes = (uint64_t)b->h_tracks[t].n5512 * sizeof(float);
      if (!bytes) continue;
      cudaError_t e = cudaMemcpyAsync(b->d_s5512 + b->h_tracks[t].s_off, tracks[t].pcm, bytes, cudaMemcpyHostToDevice,
                                      b->ctx->stream);
      if (e != cudaSuccess) return mnemo_fail(b->ctx, e, "mnemo_batch_upload");
      continue;
    }
    if (tracks[t].n_frames != b->h_tracks[t].n_pcm || tracks[t].channels != b->h_tracks[t].channels) {
      b->ctx->err = "mnemo_batch_upload: track shape differs from mnemo_batch_create";
      return -3;
    }
    const uint64_t bytes = tracks[t].n_frames * tracks[t].channels * 2ull;
    if (!bytes || !b->h_tracks[t].n_frames) continue;   // too small: nothing will read it
    cudaError_t e =
        cudaMemcpyAsync(b->d_pcm + b->pcm_off[t], tracks[t].pcm, bytes, cudaMemcpyHostToDevice, b->ctx->stream);
    if (e != cudaSuccess) return mnemo_fail(b->ctx, e, "mnemo_batch_upload");
  }
  return 0;
}

static void timing_collect(mnemo_batch* b) {
  if (!b->ev_pending) return;
  cudaEventSynchronize(b->ev[5]);
  for (int i = 0; i < 5; i++) {
    float ms = 0;
    cudaEventElapsedTime(&ms, b->ev[i], b->ev[i + 1]);
    b->stage_ms[i] += ms;
  }
  b->timed_runs++;
  b->ev_pending = false;
}

extern "C" int mnemo_batch_set_timing(mnemo_batch* b, int on) {
  if (!b) return -3;
  cudaSetDevice(b->ctx->device);
  if (on && !b->ev[0])
    for (int i = 0; i < 6; i++)
      if (cudaEventCreate(&b->ev[i]) != cudaSuccess) return -1;
  if (b->ev_pending) timing_collect(b);
  b->timing = on != 0;
  for (int i = 0; i < 5; i++) b->stage_ms[i] = 0;
  b->timed_runs = 0;
  return 0;
}

extern "C" int mnemo_batch_get_timing(mnemo_batch* b, double* stage_ms_avg, uint32_t* n_runs) {
  if (!b) return -3;
  cudaSetDevice(b->ctx->device);
  timing_collect(b);
  for (int i = 0; i < 5; i++) stage_ms_avg[i] = b->timed_runs ? b->stage_ms[i] / b->timed_runs : 0.0;
  if (n_runs) *n_runs = b->timed_runs;
  return 0;
}

extern "C" int mnemo_batch_upload_part(mnemo_batch* b, uint32_t track, uint64_t byte_offset, const void* src,
                                       uint64_t bytes) {
  if (!b || track >= b->n_tracks || (!src && bytes) || b->prenormalized) return -3;
  cudaSetDevice(b->ctx->device);
  const uint64_t total = b->h_tracks[track].n_pcm * b->h_tracks[track].channels * 2ull;
  if (byte_offset > total || bytes > total - byte_offset) {
    b->ctx->err = "mnemo_batch_upload_part: range outside the track";
    return -3;
  }
  if (!bytes || !b->h_tracks[track].n_frames) return 0;
  cudaError_t e = cudaMemcpyAsync(b->d_pcm + b->pcm_off[track] + byte_offset, src, bytes, cudaMemcpyHostToDevice,
                                  b->ctx->stream);
  if (e != cudaSuccess) return mnemo_fail(b->ctx, e, "mnemo_batch_upload_part");
  return 0;
}

extern "C" int mnemo_batch_upload_part_ticket(mnemo_batch* b, uint32_t track, uint64_t byte_offset, const void* src,
                                              uint64_t bytes, uint64_t* ticket) {
  if (!ticket) return -3;
  const int rc = mnemo_batch_upload_part(b, track, byte_offset, src, bytes);
  if (rc != 0) return rc;
  cudaEvent_t& ev = b->up_ev[b->up_seq % mnemo_batch::kUploadEvents];
  cudaError_t e = cudaSuccess;
  if (!ev) e = cudaEventCreateWithFlags(&ev, cudaEventDisableTiming);
  if (e == cudaSuccess) e = cudaEventRecord(ev, b->ctx->stream);
  if (e != cudaSuccess) return mnemo_fail(b->ctx, e, "mnemo_batch_upload_part_ticket");
  *ticket = b->up_seq++;
  return 0;
}

extern "C" int mnemo_batch_upload_wait(mnemo_batch* b, uint64_t ticket) {
  if (!b || ticket >= b->up_seq) return -3;
  // a reused event stands for a LATER copy on the same stream: waiting for it covers this ticket as well
  cudaSetDevice(b->ctx->device);
  cudaError_t e = cudaEventSynchronize(b->up_ev[ticket % mnemo_batch::kUploadEvents]);
  if (e != cudaSuccess) return mnemo_fail(b->ctx, e, "mnemo_batch_upload_wait");
  return 0;
}

extern "C" int mnemo_host_alloc(mnemo_ctx* ctx, void** p, uint64_t bytes) {
  if (!ctx || !p) return -3;
  cudaSetDevice(ctx->device);
  cudaError_t e = cudaHostAlloc(p, bytes ? bytes : 1, cudaHostAllocDefault);
  if (e != cudaSuccess) {
    *p = nullptr;
    return mnemo_fail(ctx, e, "mnemo_host_alloc");
  }
  return 0;
}

extern "C" void mnemo_host_free(mnemo_ctx* ctx, void* p) {
  if (!ctx || !p) return;
  cudaSetDevice(ctx->device);
  cudaFreeHost(p);
}

extern "C" int mnemo_batch_run(mnemo_batch* b) {
  if (!b) return -3;
  mnemo_ctx* ctx = b->ctx;
  cudaSetDevice(ctx->device);
  cudaStream_t st = ctx->stream;
  int launches = 0;
  const bool tm = b->timing;
  if (tm) {
    timing_collect(b);   // previous run's events, if any
    cudaEventRecord(b->ev[0], st);
  }
  if (!b->prenormalized) {
    const uint32_t k0_tiles = b->k0_tile_prefix[b->n_tracks];
    launch_k0_resample(b->d_tracks, b->d_k0_prefix, b->n_tracks, k0_tiles, ctx->d_tables, b->d_s5512, st);
    launches += k0_tiles ? 1 : 0;
  }
  if (tm) cudaEventRecord(b->ev[1], st);
  if (!b->prenormalized) {
    launch_sumsq(b->d_tracks, b->n_tracks, b->n_chunks, b->d_chunk_track, b->d_s5512, b->d_chunk_true,
                 b->d_chunk_prefix, b->d_chunk_prefix_end, b->d_recs, b->d_grecs, b->d_slice_maps, b->d_sumsq, b->d_rms, st, &launches);
    launch_normalize(b->d_tracks, b->d_chunk_track, b->n_chunks, b->d_s5512, b->d_rms, b->d_norm, st);
    launches += b->n_chunks ? 1 : 0;
  }
  if (tm) cudaEventRecord(b->ev[2], st);
  const uint32_t k1_tiles = b->k1_tile_prefix[b->n_tracks];
  if (ctx->fast_spectra)
    launch_k1_fast(b->d_tracks, b->d_k1_prefix, b->n_tracks, k1_tiles, ctx->d_tables, ctx->d_tw_fast, b->d_norm, b->d_bins,
                   ctx->sm_count, st);
  else
    launch_k1_spectral(b->d_tracks, b->d_k1_prefix, b->n_tracks, k1_tiles, ctx->d_tables, b->d_norm, b->d_bins,
                       ctx->sm_count, st);
  launches += k1_tiles ? 1 : 0;
  if (tm) cudaEventRecord(b->ev[3], st);
  launch_group_records(b->d_tracks, b->d_group_prefix, b->n_tracks, (uint32_t)b->n_groups, ctx->d_tables, b->d_bins,
                       b->d_gmax, b->d_imax, b->d_slot, b->d_grec, b->d_tap_px, ctx->h_tables.logf_variant, st, &launches);
  launch_k2_fingerprint(b->d_tracks, b->d_img_prefix, b->n_tracks, b->n_images, ctx->d_inv_perms, b->d_imax, b->d_slot,
                        b->d_grec, b->d_tap_px, b->d_sig_dense, b->d_keep, b->d_tap_raw, b->d_tap_img, b->d_tap_haar, st);
  launches += b->n_images ? 1 : 0;
  if (tm) cudaEventRecord(b->ev[4], st);
  launch_compact(b->d_sig_dense, b->d_keep, b->n_images, b->d_img_prefix, b->n_tracks, b->d_block_sums, b->d_pos,
                 b->d_sig_out, b->d_n_sigs, st, &launches);
  if (tm) {
    cudaEventRecord(b->ev[5], st);
    b->ev_pending = true;
  }
  b->launches_per_run = (uint32_t)launches;
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) return mnemo_fail(ctx, e, "mnemo_batch_run");
  return 0;
}

extern "C" int mnemo_batch_sync(mnemo_batch* b) {
  if (!b) return -3;
  cudaSetDevice(b->ctx->device);
  cudaError_t e = cudaStreamSynchronize(b->ctx->stream);
  if (e != cudaSuccess) return mnemo_fail(b->ctx, e, "mnemo_batch_sync");
  return 0;
}

extern "C" uint64_t mnemo_batch_max_signatures(const mnemo_batch* b) { return b ? b->n_images : 0; }
extern "C" uint32_t mnemo_batch_kernel_launches(const mnemo_batch* b) { return b ? b->launches_per_run : 0; }

extern "C" int mnemo_batch_download(mnemo_batch* b, uint8_t* sigs_out, uint64_t cap, uint32_t* n_sigs,
                                    int32_t* status) {
  if (!b) return -3;
  mnemo_ctx* ctx = b->ctx;
  cudaSetDevice(ctx->device);
  cudaStream_t st = ctx->stream;
  std::vector<uint32_t> counts(b->n_tracks);
  cudaError_t e = cudaSuccess;
  // Signatures are already compacted on the device.  Copy optimistically the upper bound that
  // fits (usually every image survives), together with the counts: one synchronisation.
  const uint64_t n_copy = b->n_images < cap ? b->n_images : cap;
  if (n_copy && sigs_out)
    e = cudaMemcpyAsync(sigs_out, b->d_sig_out, n_copy * kSigLen, cudaMemcpyDeviceToHost, st);
  if (e == cudaSuccess && b->n_tracks)
    e = cudaMemcpyAsync(counts.data(), b->d_n_sigs, 4ull * b->n_tracks, cudaMemcpyDeviceToHost, st);
  if (e == cudaSuccess) e = cudaStreamSynchronize(st);
  if (e != cudaSuccess) return mnemo_fail(ctx, e, "mnemo_batch_download");
  uint64_t total = 0;
  for (uint32_t t = 0; t < b->n_tracks; t++) {
    if (n_sigs) n_sigs[t] = counts[t];
    if (status) status[t] = b->h_tracks[t].n_frames ? 0 : -5;   // FILE_TOO_SMALL
    total += counts[t];
  }
  if (total > cap) {
    ctx->err = "mnemo_batch_download: output capacity too small";
    return -1;
  }
  return 0;
}

extern "C" int mnemo_batch_enable_taps(mnemo_batch* b, int on) {
  if (!b) return -3;
  cudaSetDevice(b->ctx->device);
  cudaStreamSynchronize(b->ctx->stream);
  if (on && !b->d_tap_raw) {
    cudaError_t e = dalloc(&b->d_tap_raw, b->n_images * kRawBytes);
    if (e == cudaSuccess) e = dalloc(&b->d_tap_img, b->n_images * kImgElems);
    if (e == cudaSuccess) e = dalloc(&b->d_tap_haar, b->n_images * kImgElems);
    if (e == cudaSuccess) e = dalloc(&b->d_tap_px, b->n_groups * 16 * 256);
    if (e != cudaSuccess) return mnemo_fail(b->ctx, e, "mnemo_batch_enable_taps");
  } else if (!on && b->d_tap_raw) {
    cudaFree(b->d_tap_raw);
    cudaFree(b->d_tap_img);
    cudaFree(b->d_tap_haar);
    cudaFree(b->d_tap_px);
    b->d_tap_px = nullptr;
    b->d_tap_raw = nullptr;
    b->d_tap_img = b->d_tap_haar = nullptr;
  }
  return 0;
}

extern "C" int64_t mnemo_batch_read(mnemo_batch* b, int what, uint32_t track, void* dst, uint64_t dst_bytes) {
  if (!b || track >= b->n_tracks) return -3;
  cudaSetDevice(b->ctx->device);
  const TrackDev& d = b->h_tracks[track];
  const void* src = nullptr;
  uint64_t bytes = 0;
  switch (what) {
    case 0: src = b->d_s5512 + d.s_off; bytes = 4ull * d.n5512; break;
    case 1: src = b->d_rms + track; bytes = 4; break;
    case 2: src = b->d_bins + d.frame_off * kBands; bytes = 4ull * d.n_frames * kBands; break;
    case 3: src = b->d_keep + d.img_off; bytes = d.n_images; break;
    case 4: src = b->d_sig_dense + d.img_off * kSigLen; bytes = (uint64_t)d.n_images * kSigLen; break;
    case 5: src = b->d_tap_raw ? b->d_tap_raw + d.img_off * kRawBytes : nullptr; bytes = (uint64_t)d.n_images * kRawBytes; break;
    case 6: src = b->d_tap_img ? b->d_tap_img + d.img_off * kImgElems : nullptr; bytes = 4ull * d.n_images * kImgElems; break;
    case 7: src = b->d_tap_haar ? b->d_tap_haar + d.img_off * kImgElems : nullptr; bytes = 4ull * d.n_images * kImgElems; break;
    case 8: src = b->d_sumsq + track; bytes = 4; break;
    case 9: src = b->d_slot + d.img_off * 16; bytes = (uint64_t)d.n_images * 16; break;
    default: return -3;
  }
  if (!src && bytes) return -3;
  if (bytes > dst_bytes) return -1;
  cudaError_t e = cudaStreamSynchronize(b->ctx->stream);
  if (e == cudaSuccess && bytes) e = cudaMemcpy(dst, src, bytes, cudaMemcpyDeviceToHost);
  if (e != cudaSuccess) return mnemo_fail(b->ctx, e, "mnemo_batch_read");
  return (int64_t)bytes;
}

extern "C" int mnemo_index_pcm_batch(mnemo_ctx* ctx, const mnemo_track* tracks, uint32_t n_tracks, uint8_t* sigs_out,
                                     uint64_t cap, uint32_t* n_sigs, int32_t* status) {
  mnemo_batch* b = nullptr;
  int rc = mnemo_batch_create(ctx, tracks, n_tracks, &b);
  if (rc != 0) return rc;
  rc = mnemo_batch_upload(b, tracks);
  if (rc == 0) rc = mnemo_batch_run(b);
  if (rc == 0) rc = mnemo_batch_download(b, sigs_out, cap, n_sigs, status);
  mnemo_batch_destroy(b);
  return rc;
}

// generate_fingerprint_from_samples (fingerprinting.c:81-109).  Its caller is a live loop that fingerprints a window of
// the same length again and again (ears/main.m:62-93), so the batch of the last length stays alive in the context
// together with pinned in/out buffers and ONE captured CUDA graph holding the copy in, the ~15 kernels and the copy
// out: a call is a memcpy into pinned memory, one graph launch, one synchronisation.
extern "C" int mnemo_index_samples(mnemo_ctx* ctx, const float* samples, uint32_t n, uint8_t* sigs_out, uint64_t cap,
                                   uint32_t* n_sigs) {
  if (n_sigs) *n_sigs = 0;
  if (!ctx || !samples) return -3;
  if (n < (uint32_t)kFrame || mnemo_frames(n) == 0) return -5;   // fingerprinting.c:82, spectralimages.c:128
  std::lock_guard<std::mutex> lock(ctx->live_mu);
  cudaSetDevice(ctx->device);
  mnemo_ctx::Live& lv = ctx->live;
  if (!lv.own) {   // the live path runs on a context (stream) of its own, under live_mu: a stream that is being captured
                   // must not receive work from any other thread, and this context's stream is shared
    int rc = mnemo_create_ex(ctx->device, nullptr, ctx->fast_spectra ? MNEMO_FAST_SPECTRA : 0u, &lv.own);
    if (rc != 0) {
      ctx->err = "mnemo_index_samples: cannot create the live path's context";
      return rc;
    }
  }
  mnemo_ctx* lc = lv.own;
  cudaStream_t st = lc->stream;
  cudaError_t e = cudaSuccess;
  if (lv.n != n || !lv.batch) {
    live_release(ctx, /*keep_context=*/true);
    mnemo_track tr;
    tr.pcm = nullptr;
    tr.n_frames = n;
    tr.channels = 1;
    tr.reserved = 0;
    int rc = batch_create_common(lc, &tr, 1, true, &lv.batch, /*share_arena=*/false);
    if (rc != 0) {
      ctx->err = lc->err;
      return rc;
    }
    mnemo_batch* b = lv.batch;
    const uint64_t out_bytes = b->n_images * kSigLen + 16;
    e = cudaHostAlloc((void**)&lv.pin_in, (uint64_t)n * sizeof(float), cudaHostAllocDefault);
    if (e == cudaSuccess) e = cudaHostAlloc((void**)&lv.pin_out, out_bytes, cudaHostAllocDefault);
    cudaGraph_t graph = nullptr;
    if (e == cudaSuccess) e = cudaStreamBeginCapture(st, cudaStreamCaptureModeThreadLocal);
    if (e == cudaSuccess) {
      tr.pcm = lv.pin_in;
      rc = mnemo_batch_upload(b, &tr);
      if (rc == 0) rc = mnemo_batch_run(b);
      cudaError_t e2 = cudaMemcpyAsync(lv.pin_out, b->d_sig_out, b->n_images * kSigLen, cudaMemcpyDeviceToHost, st);
      if (e2 == cudaSuccess)
        e2 = cudaMemcpyAsync(lv.pin_out + b->n_images * kSigLen, b->d_n_sigs, 4, cudaMemcpyDeviceToHost, st);
      e = cudaStreamEndCapture(st, &graph);
      if (e == cudaSuccess && (rc != 0 || e2 != cudaSuccess)) e = e2 != cudaSuccess ? e2 : cudaErrorUnknown;
    }
    if (e == cudaSuccess) e = cudaGraphInstantiate(&lv.exec, graph, 0);
    if (graph) cudaGraphDestroy(graph);
    if (!lv.pin_in || !lv.pin_out) {
      live_release(ctx, true);
      return mnemo_fail(ctx, e, "mnemo_index_samples (pinned buffers)");
    }
    if (e != cudaSuccess) {   // not capturable here: the same operations are issued one by one on every call
      cudaGetLastError();
      lv.exec = nullptr;
      e = cudaSuccess;
    }
    lv.n = n;
  }
  mnemo_batch* b = lv.batch;
  memcpy(lv.pin_in, samples, (uint64_t)n * sizeof(float));
  if (lv.exec) {
    e = cudaGraphLaunch(lv.exec, st);
  } else {
    mnemo_track tr;
    tr.pcm = lv.pin_in;
    tr.n_frames = n;
    tr.channels = 1;
    tr.reserved = 0;
    int rc = mnemo_batch_upload(b, &tr);
    if (rc == 0) rc = mnemo_batch_run(b);
    if (rc != 0) {
      ctx->err = lc->err;
      return rc;
    }
    e = cudaMemcpyAsync(lv.pin_out, b->d_sig_out, b->n_images * kSigLen, cudaMemcpyDeviceToHost, st);
    if (e == cudaSuccess) e = cudaMemcpyAsync(lv.pin_out + b->n_images * kSigLen, b->d_n_sigs, 4, cudaMemcpyDeviceToHost, st);
  }
  if (e == cudaSuccess) e = cudaStreamSynchronize(st);
  if (e != cudaSuccess) {
    live_release(ctx, true);
    return mnemo_fail(ctx, e, "mnemo_index_samples");
  }
  uint32_t cnt = 0;
  memcpy(&cnt, lv.pin_out + b->n_images * kSigLen, 4);
  if (cnt > cap) {
    ctx->err = "mnemo_index_samples: output capacity too small";
    return -1;
  }
  if (cnt && sigs_out) memcpy(sigs_out, lv.pin_out, (uint64_t)cnt * kSigLen);
  if (n_sigs) *n_sigs = cnt;
  return 0;
}

extern "C" int mnemo_index_samples_graph(const mnemo_ctx* ctx) { return ctx && ctx->live.exec ? 1 : 0; }

// ------------------------------------------------------------------ self tests ----

extern "C" int64_t mnemo_selftest_logf(mnemo_ctx* ctx, uint32_t lo_bits, uint32_t hi_bits) {
  if (!ctx || hi_bits < lo_bits) return -3;
  cudaSetDevice(ctx->device);
  const uint32_t kChunk = 1u << 22;
  float* d_out = nullptr;
  if (cudaMalloc(&d_out, sizeof(float) * kChunk) != cudaSuccess) return -1;
  std::vector<float> h(kChunk);
  int64_t bad = 0;
  for (uint64_t lo = lo_bits; lo <= hi_bits; lo += kChunk) {
    const uint32_t n = (uint32_t)((hi_bits - lo + 1 < kChunk) ? (hi_bits - lo + 1) : kChunk);
    launch_logf_range((uint32_t)lo, n, ctx->h_tables.logf_variant, d_out, ctx->stream);
    cudaError_t e = cudaMemcpyAsync(h.data(), d_out, sizeof(float) * n, cudaMemcpyDeviceToHost, ctx->stream);
    if (e == cudaSuccess) e = cudaStreamSynchronize(ctx->stream);
    if (e != cudaSuccess) {
      cudaFree(d_out);
      return mnemo_fail(ctx, e, "mnemo_selftest_logf");
    }
    for (uint32_t i = 0; i < n; i++) {
      const uint32_t bits = (uint32_t)lo + i;
      float x;
      memcpy(&x, &bits, 4);
      const float ref = logf(x);
      if (memcmp(&ref, &h[i], 4) != 0) bad++;
    }
  }
  cudaFree(d_out);
  return bad;
}

extern "C" int64_t mnemo_selftest_div(mnemo_ctx* ctx, int which, uint32_t lo_bits, uint64_t n) {
  if (!ctx || which < 0 || which > 4) return -3;
  cudaSetDevice(ctx->device);
  unsigned long long* d_bad = nullptr;
  if (cudaMalloc(&d_bad, 16) != cudaSuccess) return -1;
  cudaMemsetAsync(d_bad, 0, 8, ctx->stream);
  cudaMemsetAsync(d_bad + 1, 0xff, 8, ctx->stream);
  const uint64_t kChunk = 1ull << 30;
  for (uint64_t off = 0; off < n; off += kChunk)
    launch_div_test(which, (uint32_t)(lo_bits + off), (n - off < kChunk) ? (n - off) : kChunk, d_bad, ctx->h_tables.log256,
                    ctx->h_tables.inv_log256, ctx->stream);
  unsigned long long bad[2] = {0, 0};
  cudaError_t e = cudaMemcpyAsync(bad, d_bad, 16, cudaMemcpyDeviceToHost, ctx->stream);
  if (e == cudaSuccess) e = cudaStreamSynchronize(ctx->stream);
  cudaFree(d_bad);
  if (e != cudaSuccess) return mnemo_fail(ctx, e, "mnemo_selftest_div");
  if (bad[0]) {
    char buf[96];
    snprintf(buf, sizeof buf, "selftest_div(%d): first mismatch at input 0x%llx", which, bad[1]);
    ctx->err = buf;
  }
  return (int64_t)bad[0];
}
