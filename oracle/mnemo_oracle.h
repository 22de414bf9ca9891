/*
 * mnemo_oracle.h — CPU restatement of mnemophonix's fingerprint-and-match hot path.
 *
 * TEST INFRASTRUCTURE ONLY.  Nothing in the product (mnemophonix_b200/, include/) may
 * include, link or call this.  Only tests/, __graft_entry__.smoke() and bench.py's
 * cpu_baseline leg use it, and only as the checker.
 *
 * Parity status: PINNED.  Every stage below is checked bit-for-bit against the unmodified
 * reference compiled into oracle/_ref/ (tests/test_oracle_vs_ref.py) and against the
 * committed golden vectors in tests/golden/ that were generated from that reference
 * (tools/gen_golden.py).  The reference itself ships no tests or fixtures;
 * its only known-answer pin, README.md:52-56 (31444 samples -> 42 images), is tested too.
 *
 * Third-party arithmetic that is not under /root/reference: glibc 2.39 libm
 * (sinf cosf sqrtf logf log2f powf roundf fabsf) and libc qsort (stable merge sort).
 * The oracle calls the same libm/libc on the same host, as the reference does.
 */
#ifndef MNEMO_ORACLE_H
#define MNEMO_ORACLE_H
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define MO_FRAME 2048
#define MO_HOP 64
#define MO_BANDS 32
#define MO_IMG_W 128
#define MO_IMG_STEP 8
#define MO_IMG_ELEMS (MO_IMG_W * MO_BANDS)
#define MO_TOPK 200
#define MO_RAW_BYTES 1024
#define MO_SIG_LEN 100
#define MO_PERM_LEN 255
#define MO_TABLES 25

/* host tables (same libm calls as the reference) */
void mo_taps(float h[31]);
void mo_hann(float w[MO_FRAME]);
void mo_twiddles(float wr[1024], float wi[1024]);
void mo_band_edges(uint16_t e[MO_BANDS + 1]);
const uint16_t* mo_permutations(void);

/* index-side stages */
void mo_downmix(const int16_t* pcm, uint64_t n_frames, int channels, float* out44);
void mo_resample(const float* in44, uint32_t n, float* out /* n/8 */);
float mo_normalize(float* s, uint32_t n); /* returns the clamped rms */
void mo_fft(const float* x, float* re, float* im);
void mo_frame_bins(const float* frame_samples, float bins[MO_BANDS]);
void mo_scale_image(const float* bins_window /*4096*/, float* img /*4096*/);
void mo_haar(float* img);
void mo_rawfp(const float* haar_img, uint8_t bits[MO_RAW_BYTES], int* is_silence);
int mo_minhash(const uint8_t bits[MO_RAW_BYTES], uint8_t sig[MO_SIG_LEN]);

uint32_t mo_n_frames(uint32_t n_samples);
uint32_t mo_n_images(uint32_t n_frames);

/* Whole path.  Returns 0 or a reference error code (errors.h:5-29).  sigs_out must hold
 * mo_n_images(...) * 100 bytes; *n_sigs is the compacted count.  Optional taps receive
 * intermediates (may be NULL): images (scaled), haar, raw bits, silence flags. */
int mo_fingerprint_samples(const float* s5512, uint32_t n, uint8_t* sigs_out, uint32_t* n_sigs,
                           float* images, float* haar, uint8_t* raw, uint8_t* silence);
int mo_fingerprint_pcm(const int16_t* pcm, uint64_t n_frames, int channels, uint8_t* sigs_out,
                       uint32_t* n_sigs, float* s5512_out, float* rms_out);

/* glibc 2.39 logf (sysdeps/ieee754/flt-32/e_logf.c) restated for x in [2^-126, inf):
 * variant 0 = build without FMA contraction, 1 = the -mfma ifunc variant. */
float mo_logf_glibc(float x, int variant);
uint64_t mo_logf_check(uint32_t lo_bits, uint32_t hi_bits, int variant);

/* search side: database = concatenated signatures + per-entry counts */
typedef struct mo_db {
  uint32_t n_entries;
  const uint32_t* entry_start; /* n_entries+1 prefix offsets into sigs */
  const uint8_t* sigs;         /* total*100 */
} mo_db;
typedef struct mo_score {
  int32_t entry;
  float score;
  int32_t n_matches;
} mo_score;
/* returns matched entry or -7 (NO_MATCH_FOUND).  scores_out (n_entries, UNSORTED i.e. per
 * entry index) and stats (L = probed nodes, D = deep checks) may be NULL. */
int mo_search(const mo_db* db, const uint8_t* query_sigs, uint32_t n_query, mo_score* scores_out,
              uint64_t* L_out, uint64_t* D_out);
int mo_search_shard(const mo_db* shard, uint32_t table_size, const uint8_t* query_sigs, uint32_t n_query,
                    mo_score* scores_out, uint32_t* last_out /* n_query x {has, entry, sig, hits, score} */);
/* final ranking only (search.c:170-193) on a dense per-entry score array */
int mo_select(const mo_score* per_entry, uint32_t n_entries);
/* get_matches order restated (lsh.c:89-112): returns count, writes (entry,sig) pairs */
uint32_t mo_get_matches(const mo_db* db, const uint8_t* sig, uint32_t* pairs_out, uint32_t cap);

#ifdef __cplusplus
}
#endif
#endif
