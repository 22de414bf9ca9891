/*
 * mnemo_oracle.c — CPU restatement of mnemophonix's fingerprint-and-match hot path.
 * TEST INFRASTRUCTURE ONLY (see mnemo_oracle.h).  Parity status: PINNED against the
 * unmodified reference (oracle/_ref) and tests/golden/.
 *
 * Written from the arithmetic specification in SURVEY.md App. A; every function cites
 * the reference lines whose behaviour it restates.  Compile with -ffp-contract=off: the
 * reference is built for baseline x86-64 (no FMA contraction), and C's float/double
 * promotion rules are part of the behaviour.
 */
#include "mnemo_oracle.h"

#include <math.h>
#include <stdlib.h>
#include <string.h>

static const uint16_t k_perm[MO_SIG_LEN * MO_PERM_LEN] = {
#include "permutations_table.inc"
};

const uint16_t* mo_permutations(void) { return k_perm; }

/* ------------------------------------------------------------------ host tables ---- */

/* resample.c:12-35 — 31-tap Blackman-windowed sinc, cutoff 1/8. */
void mo_taps(float h[31]) {
  for (int x = -15; x <= 15; x++) {
    if (x == 0) {
      h[15] = 0.125f;
      continue;
    }
    float xs = (float)(x * 0.125);                 /* sinc(float) parameter */
    double px = M_PI * (double)xs;
    float sc = (float)((double)sinf((float)px) / px);
    float d = (float)x - 15;                       /* blackman_window(float): (x - 15) in f32 */
    double a1 = 2 * M_PI * (double)d / 30;
    double a2 = 4 * M_PI * (double)d / 30;
    float bw = (float)(0.42 - 0.5 * (double)cosf((float)a1) + 0.08 * (double)cosf((float)a2));
    h[x + 15] = (float)(0.125 * (double)sc * (double)bw);
  }
}

/* hannwindow.c:8-12 — symmetric Hann, denominator N-1, "1 - cosf" evaluated in f32. */
void mo_hann(float w[MO_FRAME]) {
  for (unsigned i = 0; i < MO_FRAME; i++) {
    float c = cosf((float)(2 * M_PI * i / (MO_FRAME - 1)));
    float one_minus = 1 - c;
    w[i] = (float)(0.5 * (double)one_minus);
  }
}

/* fft.c:47-51 — twiddle of stage l, index k is cosf/sinf((float)(-2.0*k*pi/l)); the
 * l = 2048 row reproduces every stage (k*2048/l) because the scalings are powers of 2. */
void mo_twiddles(float wr[1024], float wi[1024]) {
  for (int k = 0; k < 1024; k++) {
    float kth = (float)(-2.0 * k * M_PI / 2048);
    wr[k] = cosf(kth);
    wi[k] = sinf(kth);
  }
}

/* logbins.c:20-55 — 33 log-spaced FFT indices between 318 Hz and 2000 Hz. */
void mo_band_edges(uint16_t e[MO_BANDS + 1]) {
  float lo = log2f(318);
  float hi = log2f(2000);
  float step = (hi - lo) / MO_BANDS;
  float cur = lo;
  for (int i = 0; i <= MO_BANDS; i++) {
    float f = powf(2, cur);
    cur += step;
    int idx = (int)roundf((float)(1024.0 * (double)f / 2756.0));
    if (idx < 1) idx = 1;
    if (idx > 1024) idx = 1024;
    e[i] = (uint16_t)idx;
  }
}

/* ---------------------------------------------------------------- pre-processing ---- */

/* wav.c:358-375 — int sum of the channels, f32 divide by channel count, f64 divide by 32767. */
void mo_downmix(const int16_t* pcm, uint64_t n_frames, int channels, float* out44) {
  for (uint64_t i = 0; i < n_frames; i++) {
    int sum = 0;
    for (int c = 0; c < channels; c++) sum += pcm[i * (uint64_t)channels + c];
    float mean = (float)sum / (float)channels;
    out44[i] = (float)((double)mean / 32767.0);
  }
}

/* resample.c:38-62 — causal 31-tap dot product every 8th sample, tail truncated. */
void mo_resample(const float* in44, uint32_t n, float* out) {
  float h[31];
  mo_taps(h);
  uint32_t n_out = n / 8;
  for (uint32_t i = 0; i < n_out; i++) {
    uint32_t base = i * 8;
    float acc = 0;
    for (uint32_t j = 0; j < 31 && base + j < n; j++) {
      float p = in44[base + j] * h[j];
      acc = acc + p;
    }
    out[i] = acc;
  }
}

/* audionormalizer.c:5-32 — sequential f32 sum of squares, rms*10 clamped to [0.1, 3]. */
float mo_normalize(float* s, uint32_t n) {
  float ss = 0;
  for (uint32_t i = 0; i < n; i++) {
    float sq = s[i] * s[i];
    ss = ss + sq;
  }
  float rms = (float)((double)sqrtf(ss / (float)n) * 10.0);
  if ((double)rms < 0.1) {
    rms = (float)0.1;
  } else if ((double)rms > 3.0) {
    rms = (float)3.0;
  }
  for (uint32_t i = 0; i < n; i++) {
    float v = s[i] / rms;
    if ((double)v < -1.0) v = -1.0f;
    else if ((double)v > 1.0) v = 1.0f;
    s[i] = v;
  }
  return rms;
}

/* ---------------------------------------------------------------------- spectral ---- */

static unsigned rev11(unsigned v) {
  unsigned r = 0;
  for (int b = 0; b < 11; b++)
    if (v & (1u << b)) r |= 1u << (10 - b);
  return r;
}

/* fft.c:34-86 — radix-2 DIT on bit-reversed input, twiddles via cosf/sinf per (l,k). */
void mo_fft(const float* x, float* re, float* im) {
  for (unsigned p = 0; p < MO_FRAME; p++) {
    re[p] = x[rev11(p)];
    im[p] = 0.0f;
  }
  for (int l = 2; l <= MO_FRAME; l *= 2) {
    int half = l / 2;
    for (int k = 0; k < half; k++) {
      float ang = (float)(-2.0 * k * M_PI / l);
      float wr = cosf(ang);
      float wi = sinf(ang);
      for (int base = 0; base < MO_FRAME; base += l) {
        int a = base + k, b = a + half;
        float m1 = wr * re[b], m2 = wi * im[b];
        float m3 = wr * im[b], m4 = wi * re[b];
        float tr = m1 - m2;
        float ti = m3 + m4;
        float ar = re[a], ai = im[a];
        re[b] = ar - tr;
        im[b] = ai - ti;
        re[a] = ar + tr;
        im[a] = ai + ti;
      }
    }
  }
}

/* spectralimages.c:93-101 + logbins.c:58-76 — Hann, FFT, mean power in 32 log bands. */
void mo_frame_bins(const float* frame_samples, float bins[MO_BANDS]) {
  static float w[MO_FRAME];
  static uint16_t e[MO_BANDS + 1];
  static int ready = 0;
  if (!ready) {
    mo_hann(w);
    mo_band_edges(e);
    ready = 1;
  }
  float x[MO_FRAME], re[MO_FRAME], im[MO_FRAME];
  for (int j = 0; j < MO_FRAME; j++) x[j] = frame_samples[j] * w[j];
  mo_fft(x, re, im);
  for (int b = 0; b < MO_BANDS; b++) {
    float sum = 0;
    for (unsigned j = e[b]; j < e[b + 1]; j++) {
      float r = (float)((double)re[j] / 1024.0);
      float q = (float)((double)im[j] / 1024.0);
      float rr = r * r, qq = q * q;
      float pw = rr + qq;
      sum = sum + pw;
    }
    bins[b] = sum / (float)(unsigned)(e[b + 1] - e[b]);
  }
}

uint32_t mo_n_frames(uint32_t n_samples) { return (uint32_t)(1 + (((int)n_samples - MO_FRAME) / MO_HOP)); }
uint32_t mo_n_images(uint32_t n_frames) { return (uint32_t)(1 + (((int)n_frames - MO_IMG_W) / MO_IMG_STEP)); }

/* spectralimages.c:52-77 — running max from element 0, log(1 + 255 v/max)/log 256. */
void mo_scale_image(const float* win, float* img) {
  float mx = win[0];
  for (int i = 1; i < MO_IMG_ELEMS; i++)
    if (win[i] > mx) mx = win[i];
  float den = logf((float)256.0);
  for (int i = 0; i < MO_IMG_ELEMS; i++) {
    float sc = (float)(255.0 * (double)win[i] / (double)mx);
    if ((double)sc > 255.0) sc = (float)255.0;
    float arg = 1 + sc;
    img[i] = logf(arg) / den;
  }
}

/* haar.c:23-42 — full 1-D decomposition; (a±b) in f32, the divide by sqrt(2) in f64. */
static void haar_1d(float* d, int n) {
  float t[MO_IMG_W];
  while (n > 1) {
    n /= 2;
    for (int i = 0; i < n; i++) {
      float s = d[2 * i] + d[2 * i + 1];
      float f = d[2 * i] - d[2 * i + 1];
      t[i] = (float)((double)s / M_SQRT2);
      t[i + n] = (float)((double)f / M_SQRT2);
    }
    memcpy(d, t, sizeof(float) * 2 * (size_t)n);
  }
}

/* haar.c:48-73 — time axis (stride 32) for each band, then band axis for each frame. */
void mo_haar(float* img) {
  float line[MO_IMG_W];
  for (int b = 0; b < MO_BANDS; b++) {
    for (int t = 0; t < MO_IMG_W; t++) line[t] = img[t * MO_BANDS + b];
    haar_1d(line, MO_IMG_W);
    for (int t = 0; t < MO_IMG_W; t++) img[t * MO_BANDS + b] = line[t];
  }
  for (int t = 0; t < MO_IMG_W; t++) haar_1d(img + t * MO_BANDS, MO_BANDS);
}

typedef struct {
  float c;
  unsigned idx;
} coef_t;

static int cmp_absdesc(const void* pa, const void* pb) {
  float a = fabsf(((const coef_t*)pa)->c), b = fabsf(((const coef_t*)pb)->c);
  if (a > b) return -1;
  if (a < b) return 1;
  return 0;
}

/* rawfingerprints.c:43-100 — libc qsort by |c| desc (stable on glibc), first 200 to
 * tri-state bits, silence when fewer than 10 of them exceed 1.0 in magnitude. */
void mo_rawfp(const float* haar_img, uint8_t bits[MO_RAW_BYTES], int* is_silence) {
  coef_t v[MO_IMG_ELEMS];
  for (unsigned i = 0; i < MO_IMG_ELEMS; i++) {
    v[i].c = haar_img[i];
    v[i].idx = i;
  }
  qsort(v, MO_IMG_ELEMS, sizeof(coef_t), cmp_absdesc);
  memset(bits, 0, MO_RAW_BYTES);
  int strong = 0;
  for (int i = 0; i < MO_TOPK; i++) {
    double c = (double)v[i].c;
    if (c > 0.001) {
      unsigned pos = 2 * v[i].idx;
      bits[pos >> 3] |= (uint8_t)(1u << (pos & 7));
    } else if (c < -0.001) {
      unsigned pos = 2 * v[i].idx + 1;
      bits[pos >> 3] |= (uint8_t)(1u << (pos & 7));
    }
    if ((double)fabsf(v[i].c) > 1.0) strong++;
  }
  *is_silence = strong < 10;
}

/* minhash.c:13-28 — first set bit along each of the 100 permutations, 255 if none. */
int mo_minhash(const uint8_t bits[MO_RAW_BYTES], uint8_t sig[MO_SIG_LEN]) {
  int any = 0;
  for (int p = 0; p < MO_SIG_LEN; p++) {
    const uint16_t* row = k_perm + p * MO_PERM_LEN;
    int found = MO_PERM_LEN;
    for (int j = 0; j < MO_PERM_LEN; j++) {
      unsigned pos = row[j];
      if (bits[pos >> 3] & (1u << (pos & 7))) {
        found = j;
        break;
      }
    }
    sig[p] = (uint8_t)found;
    any |= found != MO_PERM_LEN;
  }
  return any;
}

/* fingerprinting.c:81-109 + spectralimages.c:126-221 + minhash.c:31-54. */
int mo_fingerprint_samples(const float* s, uint32_t n, uint8_t* sigs_out, uint32_t* n_sigs,
                           float* images, float* haar, uint8_t* raw, uint8_t* silence) {
  *n_sigs = 0;
  if (n < MO_FRAME) return -5;
  uint32_t nf = mo_n_frames(n);
  if (nf < MO_IMG_W) return -5;
  uint32_t ni = mo_n_images(nf);
  float* bins = (float*)malloc((size_t)nf * MO_BANDS * sizeof(float));
  if (!bins) return -1;
  for (uint32_t f = 0; f < nf; f++) mo_frame_bins(s + (size_t)f * MO_HOP, bins + (size_t)f * MO_BANDS);
  float img[MO_IMG_ELEMS];
  uint8_t bits[MO_RAW_BYTES];
  uint32_t kept = 0;
  for (uint32_t j = 0; j < ni; j++) {
    mo_scale_image(bins + (size_t)j * MO_IMG_STEP * MO_BANDS, img);
    if (images) memcpy(images + (size_t)j * MO_IMG_ELEMS, img, sizeof img);
    mo_haar(img);
    if (haar) memcpy(haar + (size_t)j * MO_IMG_ELEMS, img, sizeof img);
    int silent;
    mo_rawfp(img, bits, &silent);
    if (raw) memcpy(raw + (size_t)j * MO_RAW_BYTES, bits, MO_RAW_BYTES);
    if (silence) silence[j] = (uint8_t)silent;
    if (!silent && mo_minhash(bits, sigs_out + (size_t)kept * MO_SIG_LEN)) kept++;
  }
  free(bins);
  *n_sigs = kept;
  return 0;
}

/* fingerprinting.c:10-78 + wav.c:345-394 on in-memory PCM. */
int mo_fingerprint_pcm(const int16_t* pcm, uint64_t n_frames, int channels, uint8_t* sigs_out,
                       uint32_t* n_sigs, float* s5512_out, float* rms_out) {
  *n_sigs = 0;
  float* s44 = (float*)malloc((size_t)(n_frames ? n_frames : 1) * sizeof(float));
  uint32_t n = (uint32_t)(n_frames / 8);
  float* s = (float*)malloc((size_t)(n ? n : 1) * sizeof(float));
  if (!s44 || !s) {
    free(s44);
    free(s);
    return -1;
  }
  mo_downmix(pcm, n_frames, channels, s44);
  mo_resample(s44, (uint32_t)n_frames, s);
  free(s44);
  float rms = mo_normalize(s, n);
  if (rms_out) *rms_out = rms;
  if (s5512_out) memcpy(s5512_out, s, (size_t)n * sizeof(float));
  int rc = mo_fingerprint_samples(s, n, sigs_out, n_sigs, NULL, NULL, NULL, NULL);
  free(s);
  return rc;
}

/* ----------------------------------------------------------------- glibc logf ---- */

/* glibc 2.39 sysdeps/ieee754/flt-32/e_logf.c + e_logf_data.c (ARM optimized-routines
 * logf, N = 16 table, degree-3 polynomial, evaluated in double).  Constants read out of
 * this image's libm.so.6 (.rodata 0xb7d40..0xb7e60) and equal to the published table.
 * variant 1 follows the instruction order of the -mfma ifunc variant (4 fused ops),
 * variant 0 the baseline one.  Valid for normal positive finite x. */
static const double k_logf_tab[16][2] = {
    {0x1.661ec79f8f3bep+0, -0x1.57bf7808caadep-2}, {0x1.571ed4aaf883dp+0, -0x1.2bef0a7c06ddbp-2},
    {0x1.49539f0f010bp+0, -0x1.01eae7f513a67p-2},  {0x1.3c995b0b80385p+0, -0x1.b31d8a68224e9p-3},
    {0x1.30d190c8864a5p+0, -0x1.6574f0ac07758p-3}, {0x1.25e227b0b8eap+0, -0x1.1aa2bc79c81p-3},
    {0x1.1bb4a4a1a343fp+0, -0x1.a4e76ce8c0e5ep-4}, {0x1.12358f08ae5bap+0, -0x1.1973c5a611cccp-4},
    {0x1.0953f419900a7p+0, -0x1.252f438e10c1ep-5}, {0x1p+0, 0x0p+0},
    {0x1.e608cfd9a47acp-1, 0x1.aa5aa5df25984p-5},  {0x1.ca4b31f026aap-1, 0x1.c5e53aa362eb4p-4},
    {0x1.b2036576afce6p-1, 0x1.526e57720db08p-3},  {0x1.9c2d163a1aa2dp-1, 0x1.bc2860d22477p-3},
    {0x1.886e6037841edp-1, 0x1.1058bc8a07ee1p-2},  {0x1.767dcf5534862p-1, 0x1.4043057b6ee09p-2}};
static const double k_ln2 = 0x1.62e42fefa39efp-1;
static const double k_logf_a[3] = {-0x1.00ea348b88334p-2, 0x1.5575b0be00b6ap-2, -0x1.ffffef20a4123p-2};

float mo_logf_glibc(float x, int variant) {
  uint32_t ix;
  memcpy(&ix, &x, 4);
  if (ix == 0x3f800000u) return 0.0f;
  uint32_t tmp = ix - 0x3f330000u;
  int i = (int)((tmp >> 19) & 15u);
  int k = (int32_t)tmp >> 23;
  uint32_t iz = ix - (tmp & 0xff800000u);
  float zf;
  memcpy(&zf, &iz, 4);
  double z = (double)zf;
  double invc = k_logf_tab[i][0], logc = k_logf_tab[i][1];
  if (variant) {
    double y0 = fma((double)k, k_ln2, logc);
    double r = fma(z, invc, -1.0);
    double y = fma(r, k_logf_a[1], k_logf_a[2]);
    double r2 = r * r;
    double hi = r + y0;
    y = fma(r2, k_logf_a[0], y);
    return (float)fma(r2, y, hi);
  }
  double r = z * invc - 1.0;
  double y0 = logc + (double)k * k_ln2;
  double r2 = r * r;
  double y = k_logf_a[1] * r + k_logf_a[2];
  y = k_logf_a[0] * r2 + y;
  return (float)(y * r2 + (y0 + r));
}

/* counts floats in [lo_bits, hi_bits] where the restatement differs from this host's logf */
uint64_t mo_logf_check(uint32_t lo_bits, uint32_t hi_bits, int variant) {
  uint64_t bad = 0;
  for (uint64_t b = lo_bits; b <= hi_bits; b++) {
    uint32_t u = (uint32_t)b;
    float x, a, c;
    memcpy(&x, &u, 4);
    a = logf(x);
    c = mo_logf_glibc(x, variant);
    bad += memcmp(&a, &c, 4) != 0;
  }
  return bad;
}

/* --------------------------------------------------------------------- search ---- */

/* lsh.c:49-52 — big-endian u32 of bytes 4k..4k+3. */
static uint32_t bucket_key(const uint8_t* sig, int k) {
  const uint8_t* p = sig + 4 * k;
  return ((uint32_t)p[0] << 24) | ((uint32_t)p[1] << 16) | ((uint32_t)p[2] << 8) | (uint32_t)p[3];
}

typedef struct {
  uint32_t entry, sig;
} pair_t;

static int cmp_pair(const void* a, const void* b) {
  const pair_t *x = (const pair_t*)a, *y = (const pair_t*)b;
  if (x->entry != y->entry) return x->entry < y->entry ? -1 : 1;
  if (x->sig != y->sig) return x->sig < y->sig ? -1 : 1;
  return 0;
}

/* CSR form of lsh.c:55-86: per table, signatures ordered by slot = key % size, and inside a
 * slot by ascending (entry, sig) — the reference's chains hold the same members in
 * descending order. */
typedef struct {
  uint32_t size, total;
  uint32_t* start[MO_TABLES]; /* size+1 */
  uint32_t* ids[MO_TABLES];   /* global signature ids, ascending inside a slot */
} csr_t;

static void csr_free(csr_t* c) {
  for (int k = 0; k < MO_TABLES; k++) {
    free(c->start[k]);
    free(c->ids[k]);
  }
}

static uint32_t g_size_override = 0; /* shard tests: lsh.c:61's size computed over the WHOLE database */

static int csr_build(const mo_db* db, csr_t* c) {
  memset(c, 0, sizeof *c);
  c->total = db->entry_start[db->n_entries];
  c->size = g_size_override ? g_size_override : c->total / 2; /* lsh.c:61 */
  if (c->size == 0) return -1;
  for (int k = 0; k < MO_TABLES; k++) {
    c->start[k] = (uint32_t*)calloc((size_t)c->size + 1, sizeof(uint32_t));
    c->ids[k] = (uint32_t*)malloc((size_t)c->total * sizeof(uint32_t));
    if (!c->start[k] || !c->ids[k]) return -2;
    uint32_t* st = c->start[k];
    for (uint32_t g = 0; g < c->total; g++) st[bucket_key(db->sigs + (size_t)g * MO_SIG_LEN, k) % c->size + 1]++;
    for (uint32_t s = 0; s < c->size; s++) st[s + 1] += st[s];
    uint32_t* fill = (uint32_t*)malloc((size_t)c->size * sizeof(uint32_t));
    if (!fill) return -2;
    memcpy(fill, st, (size_t)c->size * sizeof(uint32_t));
    for (uint32_t g = 0; g < c->total; g++)
      c->ids[k][fill[bucket_key(db->sigs + (size_t)g * MO_SIG_LEN, k) % c->size]++] = g;
    free(fill);
  }
  return 0;
}

static uint32_t entry_of(const mo_db* db, uint32_t g) {
  uint32_t lo = 0, hi = db->n_entries; /* last e with entry_start[e] <= g */
  while (hi - lo > 1) {
    uint32_t mid = (lo + hi) / 2;
    if (db->entry_start[mid] <= g) lo = mid;
    else hi = mid;
  }
  return lo;
}

/* search.c:65-107 — NOT a strict weak order: integer abs() of a float difference. */
static int cmp_score(const void* pa, const void* pb) {
  const mo_score *a = (const mo_score*)pa, *b = (const mo_score*)pb;
  float avg_a = a->n_matches == 0 ? 0 : a->score / (float)a->n_matches;
  float avg_b = b->n_matches == 0 ? 0 : b->score / (float)b->n_matches;
  float delta = (float)abs((int)(avg_a - avg_b));
  if (delta <= 3) {
    if (delta <= 5 && a->n_matches >= b->n_matches + 5) return -1;
    if (b->n_matches >= a->n_matches + 5) return 1;
  }
  if ((double)delta < 0.5) {
    if (a->n_matches > b->n_matches) return -1;
    if (a->n_matches < b->n_matches) return 1;
  }
  if (avg_a > avg_b) return -1;
  if (avg_b > avg_a) return 1;
  return 0;
}

/* search.c:170-193 — libc qsort of all entries, then thresholds over the first ten. */
int mo_select(const mo_score* per_entry, uint32_t n_entries) {
  mo_score* v = (mo_score*)malloc((size_t)(n_entries ? n_entries : 1) * sizeof(mo_score));
  if (!v) return -1;
  memcpy(v, per_entry, (size_t)n_entries * sizeof(mo_score));
  qsort(v, n_entries, sizeof(mo_score), cmp_score);
  int best = -7;
  float best_avg = 0;
  for (uint32_t i = 0; i < n_entries && i < 10; i++) {
    float avg = v[i].n_matches == 0 ? 0 : v[i].score / (float)v[i].n_matches;
    if ((v[i].n_matches >= 10 || (avg >= 35 && v[i].n_matches >= 5)) && avg >= 30)
      if (avg > best_avg) {
        best_avg = avg;
        best = v[i].entry;
      }
  }
  free(v);
  return best;
}

static uint32_t gather(const mo_db* db, const csr_t* c, const uint8_t* q, pair_t* out, uint32_t cap, int ref_order) {
  uint32_t n = 0;
  /* lsh.c:89-112 prepends while walking tables 0..24 and each chain head-first
   * (descending), so the returned list reads: table 24 ascending, ..., table 0 ascending. */
  for (int kk = 0; kk < MO_TABLES; kk++) {
    int k = ref_order ? MO_TABLES - 1 - kk : kk;
    uint32_t slot = bucket_key(q, k) % c->size;
    for (uint32_t p = c->start[k][slot]; p < c->start[k][slot + 1]; p++) {
      uint32_t g = c->ids[k][p];
      if (n < cap) {
        uint32_t e = entry_of(db, g);
        out[n].entry = e;
        out[n].sig = g - db->entry_start[e];
      }
      n++;
    }
  }
  return n;
}

uint32_t mo_get_matches(const mo_db* db, const uint8_t* sig, uint32_t* pairs_out, uint32_t cap) {
  csr_t c;
  if (csr_build(db, &c) != 0) {
    csr_free(&c);
    return 0;
  }
  pair_t* tmp = (pair_t*)malloc((size_t)(cap ? cap : 1) * sizeof(pair_t));
  uint32_t n = gather(db, &c, sig, tmp, cap, 1);
  for (uint32_t i = 0; i < n && i < cap; i++) {
    pairs_out[2 * i] = tmp[i].entry;
    pairs_out[2 * i + 1] = tmp[i].sig;
  }
  free(tmp);
  csr_free(&c);
  return n;
}

/* search.c:110-193.  Note the group flush exists only inside the scan loop
 * (search.c:147-165): the group with the greatest (entry, sig) is never scored. */
static int mo_search_impl(const mo_db* db, const uint8_t* qs, uint32_t nq, mo_score* scores_out, uint64_t* L_out,
                          uint64_t* D_out, uint32_t* last_out /* nq x 5: has, entry, sig, hits, score */) {
  csr_t c;
  int rc = csr_build(db, &c);
  if (rc != 0) {
    csr_free(&c);
    return rc == -1 ? -7 : -1; /* size == 0: the reference divides by zero; defined here as no match */
  }
  mo_score* sc = (mo_score*)calloc(db->n_entries ? db->n_entries : 1, sizeof(mo_score));
  for (uint32_t e = 0; e < db->n_entries; e++) sc[e].entry = (int32_t)e;
  uint64_t L = 0, D = 0;
  uint32_t cap = 1u << 16;
  pair_t* cand = (pair_t*)malloc((size_t)cap * sizeof(pair_t));
  for (uint32_t i = 0; i < nq; i++) {
    const uint8_t* q = qs + (size_t)i * MO_SIG_LEN;
    uint32_t n = gather(db, &c, q, cand, cap, 0);
    if (n > cap) {
      cap = n;
      cand = (pair_t*)realloc(cand, (size_t)cap * sizeof(pair_t));
      n = gather(db, &c, q, cand, cap, 0);
    }
    L += n;
    qsort(cand, n, sizeof(pair_t), cmp_pair);
    if (last_out) {
      uint32_t* lo = last_out + 5 * (size_t)i;
      lo[0] = n > 0;
      lo[1] = lo[2] = lo[3] = lo[4] = 0;
      if (n > 0) {
        uint32_t hits = 1;
        for (uint32_t j = n - 1; j > 0 && cmp_pair(&cand[j], &cand[j - 1]) == 0; j--) hits++;
        const uint8_t* d = db->sigs + ((size_t)db->entry_start[cand[n - 1].entry] + cand[n - 1].sig) * MO_SIG_LEN;
        unsigned same = 0;
        for (int b = 0; b < MO_SIG_LEN; b++) same += d[b] == q[b];
        lo[1] = cand[n - 1].entry;
        lo[2] = cand[n - 1].sig;
        lo[3] = hits;
        lo[4] = same;
      }
    }
    uint32_t run = 1;
    for (uint32_t j = 1; j < n; j++) {
      if (cmp_pair(&cand[j], &cand[j - 1]) == 0) {
        run++;
        continue;
      }
      if (run >= 2) {
        D++;
        const uint8_t* d = db->sigs + ((size_t)db->entry_start[cand[j - 1].entry] + cand[j - 1].sig) * MO_SIG_LEN;
        unsigned same = 0;
        for (int b = 0; b < MO_SIG_LEN; b++) same += d[b] == q[b];
        if (same >= 30) {
          sc[cand[j - 1].entry].score += (float)same;
          sc[cand[j - 1].entry].n_matches++;
        }
      }
      run = 1;
    }
  }
  free(cand);
  csr_free(&c);
  if (scores_out) memcpy(scores_out, sc, (size_t)db->n_entries * sizeof(mo_score));
  if (L_out) *L_out = L;
  if (D_out) *D_out = D;
  int best = mo_select(sc, db->n_entries);
  free(sc);
  return best;
}

int mo_search(const mo_db* db, const uint8_t* qs, uint32_t nq, mo_score* scores_out, uint64_t* L_out,
              uint64_t* D_out) {
  return mo_search_impl(db, qs, nq, scores_out, L_out, D_out, NULL);
}

/* One shard of a database split by contiguous entry ranges: `table_size` is the size of the whole
 * database's tables (lsh.c:61).  Besides the shard-local scores (its own last group dropped, as the
 * reference's loop would) it reports that last group so that a merge can restore it. */
int mo_search_shard(const mo_db* shard, uint32_t table_size, const uint8_t* qs, uint32_t nq, mo_score* scores_out,
                    uint32_t* last_out) {
  g_size_override = table_size;
  int rc = mo_search_impl(shard, qs, nq, scores_out, NULL, NULL, last_out);
  g_size_override = 0;
  return rc;
}
