// api_internal.h — private definitions behind the opaque handles of include/mnemo_cuda.h.
#pragma once
#include <mutex>
#include <string>
#include <vector>

#include "common.cuh"

struct mnemo_ctx {
  int device = 0;
  int sm_count = 0;
  cudaStream_t stream = nullptr;
  bool own_stream = false;
  mnemo::Tables h_tables;
  mnemo::Tables* d_tables = nullptr;
  uint16_t* d_inv_perms = nullptr;  // [8192][100]: position of bit b in permutation p (255 = not among the first 255)
  uint32_t logf_probes = 0, logf_matched = 0;
  bool fast_spectra = false;        // tolerance mode of K1 (mnemo_create_ex, MNEMO_FAST_SPECTRA)
  float2* d_tw_fast = nullptr;      // its twiddle tables
  std::string err;
  // Live path (mnemo_index_samples, the ears loop's once-a-second call): the batch of the last sample count is kept,
  // with pinned in/out buffers and the whole call (copy in, kernels, copy out) captured as one CUDA graph.
  struct Live {
    mnemo_ctx* own = nullptr;       // private context (own stream): nothing else can enqueue on a stream being captured
    struct mnemo_batch* batch = nullptr;
    uint32_t n = 0;
    float* pin_in = nullptr;
    uint8_t* pin_out = nullptr;     // signatures, then the count
    cudaGraphExec_t exec = nullptr;
  } live;
  std::mutex live_mu;
  // Device arena handed from batch to batch: a batch takes all its buffers from ONE allocation, and the context
  // keeps the last one for the next batch (a process that indexes file after file pays cudaMalloc / cudaFree —
  // milliseconds for gigabytes — once, not per file).  A second batch alive at the same time gets its own.
  uint8_t* arena = nullptr;
  uint64_t arena_bytes = 0;
  bool arena_busy = false;
  std::mutex arena_mu;
};

struct mnemo_batch {
  mnemo_ctx* ctx = nullptr;
  uint32_t n_tracks = 0;
  bool prenormalized = false;
  std::vector<mnemo::TrackDev> h_tracks;
  std::vector<uint32_t> k0_tile_prefix, k1_tile_prefix;
  std::vector<uint64_t> img_prefix, pcm_off;
  std::vector<uint32_t> group_prefix;        // 8-frame groups per track (k15_groups.cu)
  uint64_t n_groups = 0;
  uint64_t n_samples = 0, n_frames = 0, n_images = 0, n_chunks = 0;
  uint32_t launches_per_run = 0;
  // optional per-stage timing (CUDA events on the context's stream)
  bool timing = false;
  cudaEvent_t ev[6] = {nullptr, nullptr, nullptr, nullptr, nullptr, nullptr};
  double stage_ms[5] = {0, 0, 0, 0, 0};
  uint32_t timed_runs = 0;
  bool ev_pending = false;
  // tickets of mnemo_batch_upload_part_ticket: an event per asynchronous piece, reused round-robin
  static constexpr int kUploadEvents = 64;
  cudaEvent_t up_ev[kUploadEvents] = {};
  uint64_t up_seq = 0;
  // device buffers: carved out of `arena` (one allocation; the context's when arena_of_ctx)
  uint8_t* arena = nullptr;
  bool arena_of_ctx = false;
  uint8_t* d_pcm = nullptr;
  mnemo::TrackDev* d_tracks = nullptr;
  uint32_t *d_k0_prefix = nullptr, *d_k1_prefix = nullptr, *d_chunk_track = nullptr;
  uint64_t* d_img_prefix = nullptr;
  float* d_s5512 = nullptr;
  float* d_norm = nullptr;   // normalised samples (aliases d_s5512 for pre-normalised input)
  double *d_chunk_true = nullptr, *d_chunk_prefix = nullptr;
  double* d_chunk_prefix_end = nullptr;
  mnemo::SumChunkRec* d_recs = nullptr;
  mnemo::SumChunkRec* d_grecs = nullptr;    // composed maps per group of 32 chunks
  uint32_t* d_slice_maps = nullptr;         // per-lane maps of every chunk (replay of binade-crossing chunks)
  float *d_sumsq = nullptr, *d_rms = nullptr, *d_bins = nullptr;
  uint8_t *d_sig_dense = nullptr, *d_keep = nullptr, *d_sig_out = nullptr;
  uint32_t *d_pos = nullptr, *d_block_sums = nullptr, *d_n_sigs = nullptr;
  // group records: what consecutive images share (k15_groups.cu)
  uint32_t* d_group_prefix = nullptr;
  uint32_t* d_gmax = nullptr;     // [n_groups] largest energy of the group (float bits)
  float* d_imax = nullptr;        // [n_images] window maximum
  uint8_t* d_slot = nullptr;      // [n_images][16] record slot of each of the image's groups
  float* d_grec = nullptr;        // [n_groups][16][256] records (slots in use only)
  float* d_tap_px = nullptr;      // same shape: scaled pixels (debug taps only)
  uint8_t* d_tap_raw = nullptr;
  float *d_tap_img = nullptr, *d_tap_haar = nullptr;
};

int mnemo_fail(mnemo_ctx* ctx, cudaError_t e, const char* what);
