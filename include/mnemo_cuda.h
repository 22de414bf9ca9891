/*
 * mnemo_cuda.h — C ABI of libmnemo_cuda.so, the sm_100a implementation of mnemophonix's
 * fingerprint-and-match hot path.
 *
 * Plain C: pointers, sizes and int return codes only.  Return codes are the reference's
 * (errors.h:5-29): 0 SUCCESS, -1 MEMORY_ERROR, -2 CANNOT_READ_FILE, -3 DECODING_ERROR,
 * -4 UNSUPPORTED_WAVE_FORMAT, -5 FILE_TOO_SMALL, -6 NOT_A_WAVE_FILE, -7 NO_MATCH_FOUND;
 * CUDA failures map to -1 (allocation) or -3 (anything else) and mnemo_last_error() holds
 * the text.  There is no CPU fallback: every entry point fails if no sm_100 device is there.
 *
 * What each group replaces in the reference:
 *   mnemo_batch_* / mnemo_index_pcm_batch   read_samples' downmix loop (wav.c:358-375),
 *        resample (resample.c:47-62), normalize (audionormalizer.c:5-32),
 *        build_spectral_images (spectralimages.c:126-221), apply_Haar_transform (haar.c:85-108),
 *        build_raw_fingerprints (rawfingerprints.c:103-144), build_signatures (minhash.c:31-54),
 *        i.e. the body of generate_fingerprint (fingerprinting.c:10-78) after the WAV header.
 *   mnemo_index_samples                     generate_fingerprint_from_samples (fingerprinting.c:81-109)
 *   mnemo_db_* / mnemo_search_batch         create_hash_tables / get_matches (lsh.c:55-112) and
 *        search (search.c:110-193)
 * The reference-named drop-in functions themselves (generate_fingerprint, create_hash_tables,
 * search, ...) live in libmnemophonix.so (mnemophonix_b200/host/), which is thin C over this ABI.
 */
#ifndef MNEMO_CUDA_H
#define MNEMO_CUDA_H
#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define MNEMO_SIGNATURE_LENGTH 100

typedef struct mnemo_ctx mnemo_ctx;
typedef struct mnemo_batch mnemo_batch;
typedef struct mnemo_db mnemo_db;

/* One context per (host thread or process, device).  `stream` is a cudaStream_t to run on
 * (NULL = the context creates its own non-blocking stream).  Computes the Hann window, FIR
 * taps, FFT twiddles and band edges with the host's libm — the same calls the reference
 * makes (hannwindow.c:10, resample.c:27-35, fft.c:49-51, logbins.c:44-55) — and uploads them. */
int mnemo_create(int device, void* stream, mnemo_ctx** out);
/* The same with flags.  MNEMO_FAST_SPECTRA selects the TOLERANCE mode of the frame -> band-energy kernel: a real-input
 * 1024-point four-step FFT with fused multiply-adds and double-precision twiddles instead of the bit-exact replay of
 * fft.c:34-86.  Band energies then agree with the reference to ~1e-6 relative (spectral images within BASELINE.json's
 * 1e-4), raw fingerprints on >= 99.9 % of their bits, and signatures wherever the top-200 set agrees — not bit for
 * bit.  Never the default; mnemo_create honours the environment variable MNEMO_FAST_SPECTRA=1. */
#define MNEMO_FAST_SPECTRA 1u
int mnemo_create_ex(int device, void* stream, uint32_t flags, mnemo_ctx** out);
int mnemo_spectra_mode(const mnemo_ctx* ctx); /* 0 exact, 1 tolerance mode */
void mnemo_destroy(mnemo_ctx* ctx);
const char* mnemo_last_error(const mnemo_ctx* ctx);
/* Diagnostics: which glibc logf variant (0 baseline, 1 fma) the device code mirrors, and how
 * many of the probe inputs matched the host's logf when the context was created. */
int mnemo_logf_variant(const mnemo_ctx* ctx, uint32_t* probes, uint32_t* matched);

/* ---- sizes (App. C of SURVEY.md; spectralimages.c:37-49) ---- */
uint32_t mnemo_samples_5512(uint64_t n_pcm_frames); /* n/8 */
uint32_t mnemo_frames(uint32_t n_samples_5512);     /* 0 if too small */
uint32_t mnemo_images(uint32_t n_samples_5512);     /* 0 if too small */

/* ---- indexing ---- */
typedef struct mnemo_track {
  const void* pcm;      /* interleaved little-endian int16, host memory (may be NULL for batch_create) */
  uint64_t n_frames;    /* sample frames (data bytes / block align, wav.c:347) */
  uint32_t channels;    /* >= 1 */
  uint32_t reserved;
} mnemo_track;

/* Staged form: allocate device buffers for a set of tracks, upload PCM, run the kernels
 * (asynchronously on the context's stream), download.  A batch can be re-uploaded and re-run. */
int mnemo_batch_create(mnemo_ctx* ctx, const mnemo_track* tracks, uint32_t n_tracks, mnemo_batch** out);
void mnemo_batch_destroy(mnemo_batch* b);
int mnemo_batch_upload(mnemo_batch* b, const mnemo_track* tracks); /* H2D of every track's PCM, async */
/* Streaming ingest (replaces read_samples' per-block fread loop, wav.c:345-394, for files that are read in
 * pieces): bytes [byte_offset, byte_offset + bytes) of one track's interleaved int16 data, asynchronous on the
 * context's stream when `src` is pinned (mnemo_host_alloc).  The caller reads the next piece of the file into a
 * second staging buffer meanwhile and calls mnemo_batch_sync before reusing a buffer. */
int mnemo_batch_upload_part(mnemo_batch* b, uint32_t track, uint64_t byte_offset, const void* src, uint64_t bytes);
/* The same with a ticket: mnemo_batch_upload_wait(ticket) returns once THAT piece (and every earlier one) has left its
 * staging buffer, so a ring of pinned buffers can be refilled by reader threads while later pieces are still in flight. */
int mnemo_batch_upload_part_ticket(mnemo_batch* b, uint32_t track, uint64_t byte_offset, const void* src, uint64_t bytes,
                                   uint64_t* ticket);
int mnemo_batch_upload_wait(mnemo_batch* b, uint64_t ticket);
/* Pinned host memory for callers that do not link the CUDA runtime themselves. */
int mnemo_host_alloc(mnemo_ctx* ctx, void** p, uint64_t bytes);
void mnemo_host_free(mnemo_ctx* ctx, void* p);
int mnemo_batch_run(mnemo_batch* b);                               /* kernels only, async */
int mnemo_batch_sync(mnemo_batch* b);
uint64_t mnemo_batch_max_signatures(const mnemo_batch* b);         /* = total spectral images */
uint32_t mnemo_batch_kernel_launches(const mnemo_batch* b);        /* launches issued by one run */
/* After run (+sync is implied): signatures of all tracks, compacted and concatenated in track
 * order, into sigs_out (capacity in signatures); n_sigs[t] and status[t] per track (0 or
 * FILE_TOO_SMALL). */
int mnemo_batch_download(mnemo_batch* b, uint8_t* sigs_out, uint64_t cap_signatures, uint32_t* n_sigs,
                         int32_t* status);
/* Intermediates for stage-level parity tests.  what: 0 = 5512 Hz samples before normalisation
 * (float[n/8]), 1 = rms (float[1]), 2 = band energies (float[n_frames*32]),
 * 3 = per-image keep flags (u8[n_images]), 4 = dense signatures (u8[n_images*100]),
 * 5 = raw fingerprint bits (u8[n_images*1024], only if debug taps enabled),
 * 6 = scaled spectral images, 7 = Haar coefficients (float[n_images*4096], debug taps),
 * 8 = float32 sequential sum of squares (float[1]), 9 = record slot of each of an image's 16 frame groups
 * (u8[n_images*16], k15_groups.cu: how many distinct window maxima the images around a group have).
 * Returns bytes written or <0. */
int64_t mnemo_batch_read(mnemo_batch* b, int what, uint32_t track, void* dst, uint64_t dst_bytes);
int mnemo_batch_enable_taps(mnemo_batch* b, int on);
/* Per-stage device times (CUDA events on the context's stream), averaged over the runs since timing
 * was switched on: [0] downmix+resample, [1] sum of squares + rms, [2] frames -> band energies,
 * [3] image -> signature, [4] compaction.  Milliseconds. */
int mnemo_batch_set_timing(mnemo_batch* b, int on);
int mnemo_batch_get_timing(mnemo_batch* b, double* stage_ms_avg /* 5 */, uint32_t* n_runs);

/* One-call form with host buffers: create + upload + run + download + destroy. */
int mnemo_index_pcm_batch(mnemo_ctx* ctx, const mnemo_track* tracks, uint32_t n_tracks, uint8_t* sigs_out,
                          uint64_t cap_signatures, uint32_t* n_sigs, int32_t* status);
/* generate_fingerprint_from_samples (fingerprinting.c:81-109): samples already mono, 5512 Hz
 * and normalised.  Returns 0 / FILE_TOO_SMALL / error; *n_sigs signatures written. */
int mnemo_index_samples(mnemo_ctx* ctx, const float* samples, uint32_t n, uint8_t* sigs_out,
                        uint64_t cap_signatures, uint32_t* n_sigs);
/* 1 if the live path of the last mnemo_index_samples length runs as one captured CUDA graph, 0 if call by call. */
int mnemo_index_samples_graph(const mnemo_ctx* ctx);

/* ---- search ---- */
/* Database = signatures of all entries concatenated in entry order; entry_start has
 * n_entries+1 offsets.  `table_size` is lsh.c:61's size = total_signatures/2 over the WHOLE
 * database (pass 0 to derive it from this shard alone); entry_base is the global index of the
 * shard's first entry (database sharded across GPUs by contiguous entry ranges). */
int mnemo_db_create(mnemo_ctx* ctx, const uint8_t* sigs, const uint32_t* entry_start, uint32_t n_entries,
                    uint32_t table_size, uint32_t entry_base, mnemo_db** out);
void mnemo_db_destroy(mnemo_db* db);
uint32_t mnemo_db_table_size(const mnemo_db* db);
/* 1 if the database carries the pair index (signatures ordered by the buckets of every pair of tables: candidates that share
 * two buckets with a query signature are looked up instead of found by streaming the 25 probed buckets).  Built when
 * the bucket load calls for it — sum over all buckets of len^2 / signatures, the number of ids a query signature that
 * resembles the database meets in its 25 buckets, is 12 000 or more — and device memory allows (2400 + 128 bytes per
 * signature); the environment variable MNEMO_PAIR_INDEX=1 / 0 forces it on / off.  Results are identical either way. */
int mnemo_db_has_pair_index(const mnemo_db* db);

typedef struct mnemo_match {
  int32_t entry;      /* matched entry (global index) or -7 */
  float score;        /* sum of deep-check scores of that entry */
  int32_t n_matches;
} mnemo_match;

/* get_matches (lsh.c:89-112) for one signature: (entry, signature) pairs in the order the
 * reference's list is returned.  Returns the count (may exceed cap; only cap are written). */
int64_t mnemo_db_get_matches(mnemo_db* db, const uint8_t* sig, uint32_t* pairs_out, uint64_t cap_pairs);

/* search() (search.c:110-193) for a batch of queries.  query_start has n_queries+1 offsets into
 * query_sigs.  out[q] receives the reference's return value and that entry's (score, n_matches).
 * stats (optional, 2 values): total probed list nodes L and deep checks D. */
int mnemo_search_batch(mnemo_db* db, const uint8_t* query_sigs, const uint32_t* query_start,
                       uint32_t n_queries, mnemo_match* out, uint64_t* stats);

/* mnemo_search_batch plus, per query, the matched records among the first ten of the reference's sorted scores[] array
 * (search.c:170-175), in that order: lists_out[q][10], counts_out[q] (the entries that follow them in the reference's
 * array are the unmatched entries in index order).  What search(..., verbose) prints (search.c:178). */
struct mnemo_sel_rec;
int mnemo_search_batch_top(mnemo_db* db, const uint8_t* query_sigs, const uint32_t* query_start, uint32_t n_queries,
                           mnemo_match* out, struct mnemo_sel_rec* lists_out, uint32_t* counts_out);

/* The dense rows search.c:55-59 accumulates, for parity checks: score_out / n_matches_out are [n_queries][n_entries of this
 * shard] (host), before any cross-shard fix-up. */
int mnemo_search_scores(mnemo_db* db, const uint8_t* query_sigs, const uint32_t* query_start, uint32_t n_queries,
                        uint32_t* score_out, uint32_t* n_matches_out);

/* Record form of a sharded search (arbitrary entry ranges; diagnostic — the production protocol is the ranked one
 * below): per-query candidate records of THIS shard, to be concatenated over the shards and finished on the host with
 * mnemo_search_finish. */
typedef struct mnemo_cand {
  uint32_t query;
  int32_t entry;      /* global entry index */
  float score;
  int32_t n_matches;
} mnemo_cand;
typedef struct mnemo_lastgroup {
  uint32_t has;       /* this shard saw >=1 candidate for the query signature */
  uint32_t entry;     /* greatest (entry, sig) among this shard's candidates ... */
  uint32_t sig;
  uint32_t hits;      /* ... its bucket-hit count ... */
  uint32_t score;     /* ... and its deep-check score (0 if hits < 2) */
} mnemo_lastgroup;
int mnemo_search_shard(mnemo_db* db, const uint8_t* query_sigs, const uint32_t* query_start,
                       uint32_t n_queries, mnemo_cand* cands, uint64_t cap_cands, uint64_t* n_cands,
                       mnemo_lastgroup* last /* one per query signature */, uint64_t* stats);
int mnemo_search_finish(const mnemo_cand* cands, uint64_t n_cands, const mnemo_lastgroup* last_all_shards,
                        uint32_t n_shards, const uint32_t* query_start, uint32_t n_queries,
                        uint32_t n_entries_total, mnemo_match* out);

/* Sharded search ranked on the GPUs: THE multi-GPU protocol (search.c:170-193 needs only the first ten records of
 * glibc's merge sort over all entries; every node of that merge tree can be ranked where its entries live).  The database
 * is sharded on node boundaries of the tree: mnemo_tree_cuts fills cuts[n_shards + 1] (entry ranges of the shards) and
 * node_lo[n_nodes + 1] (entry ranges of the n_nodes = 2^ceil(log2 n_shards) nodes; shard r owns nodes
 * r * n_nodes / n_shards .. (r + 1) * n_nodes / n_shards - 1) and returns n_nodes.
 *
 * Device-resident form (everything asynchronous on the context's stream; the two all-gathers between the calls are the
 * caller's, e.g. NCCL on the same stream).  Per tile of at most mnemo_db_tile_queries() queries:
 *   mnemo_search_dense_begin_dev   probes this shard (host query signatures in, accumulators stay on the device) and
 *                                  writes d_has_out[s] = the shard holds a candidate for query signature s;
 *   -- all-gather of the flags: d_has_all[n_shards][n_query_signatures] --
 *   mnemo_search_dense_lists_dev   where a higher shard holds candidates the globally greatest (entry, signature)
 *                                  group, which search.c:147-165 never scores, lives there, so this shard's own
 *                                  greatest group is scored after all; then every owned node is ranked:
 *                                  d_lists_out[node][query][10] + d_counts_out[node][query] (device memory; a shard's
 *                                  block of mnemo_shard_record_bytes() is lists followed by counts);
 *   -- all-gather of the blocks --
 *   mnemo_select_merge_dev         merges the lists of all nodes through the top levels of the tree on the GPU, applies
 *                                  the thresholds, returns 12 bytes per query to the host (synchronises).
 * Host-pointer forms of the same steps: mnemo_search_dense_begin / _lists (has_higher[s] = some shard with greater
 * entry indices has a candidate) and mnemo_select_merge (lists of ALL nodes in node order). */
typedef struct mnemo_sel_rec {
  int32_t entry;      /* global entry index */
  float score;
  int32_t n_matches;
  float avg;          /* score / n_matches (search.c:66-67) */
} mnemo_sel_rec;
uint32_t mnemo_tree_cuts(uint32_t n_entries_total, uint32_t n_shards, uint32_t* cuts, uint32_t* node_lo, uint32_t node_cap);
uint32_t mnemo_db_tile_queries(const mnemo_db* db);   /* queries whose dense accumulators fit the budget (1 GiB) */
int mnemo_search_dense_begin_dev(mnemo_db* db, const uint8_t* query_sigs, const uint32_t* query_start, uint32_t n_queries,
                                 uint8_t* d_has_out /* device, one per query signature */);
int mnemo_search_dense_lists_dev(mnemo_db* db, const uint8_t* d_has_all /* device */, uint32_t n_shards, uint32_t shard,
                                 const uint32_t* node_bounds /* host */, uint32_t n_nodes, mnemo_sel_rec* d_lists_out,
                                 uint32_t* d_counts_out);
uint64_t mnemo_shard_record_bytes(uint32_t nodes_per_shard, uint32_t n_queries);
int mnemo_select_merge_dev(mnemo_ctx* ctx, const void* d_gathered /* device: n_shards blocks */, uint32_t n_shards,
                           uint32_t n_nodes, uint32_t nodes_per_shard, uint32_t n_queries, mnemo_match* out /* host */);
int mnemo_search_dense_stats(mnemo_db* db, uint64_t* stats /* L, D of the last probe */);
/* Diagnostics for parity tests: the accumulator row (score sum, n_matches per entry of this shard, search.c:55-59) of
 * one query of the batch held since mnemo_search_dense_begin*, after the fix-up when _lists has run. */
int mnemo_search_dense_read(mnemo_db* db, uint32_t query, uint32_t* score_out, uint32_t* n_matches_out);
int mnemo_search_dense_begin(mnemo_db* db, const uint8_t* query_sigs, const uint32_t* query_start, uint32_t n_queries,
                             uint8_t* has_out /* one per query signature */, uint64_t* stats);
int mnemo_search_dense_lists(mnemo_db* db, const uint8_t* has_higher, const uint32_t* node_bounds, uint32_t n_nodes,
                             mnemo_sel_rec* lists_out, uint32_t* counts_out);
int mnemo_select_merge(const mnemo_sel_rec* lists, const uint32_t* counts, uint32_t n_nodes, uint32_t n_queries,
                       mnemo_match* out);

/* ---- self tests of the exact-arithmetic building blocks (run on the device) ---- */
/* Compares the device logf against host logf for every float in [lo_bits, hi_bits]; returns the
 * number of mismatches (or <0).  Used to pin the glibc logf mirror exhaustively over [1, 256]. */
int64_t mnemo_selftest_logf(mnemo_ctx* ctx, uint32_t lo_bits, uint32_t hi_bits);
/* Compares the reciprocal-based exact divisions of the kernels with IEEE division on the device, for the n
 * floats whose bit patterns start at lo_bits (n = 2^32 covers every float); returns the number of mismatches.
 * which 0: x / M_SQRT2 in double (haar.c:35-36); 1: x / 32767.0 (wav.c:373) in FP32, plain and halving forms;
 * 3: x / M_SQRT2 in FP32 on its domain (0 or |x| >= 2^-103); 4: x / logf(256.0f) (spectralimages.c:57) in FP32
 * on its domain (0 or |x| >= 2^-106); 2: (255*v) / max (spectralimages.c:53) for n pseudo-random pairs. */
int64_t mnemo_selftest_div(mnemo_ctx* ctx, int which, uint32_t lo_bits, uint64_t n);

/* The ranking kernel (search.c:65-107,170-193 on the GPU) on caller-supplied per-entry rows score[q][e], n_matches[q][e]
 * (n_queries x n_entries, host memory); out[q] as mnemo_search_batch.  entry_base shifts the rows to tree positions
 * entry_base .. entry_base + n_entries - 1 (the entries below have no matches). */
int mnemo_selftest_select(mnemo_ctx* ctx, const uint32_t* score, const uint32_t* n_matches, uint32_t n_queries,
                          uint32_t n_entries, uint32_t entry_base, mnemo_match* out);

#ifdef __cplusplus
}
#endif
#endif
