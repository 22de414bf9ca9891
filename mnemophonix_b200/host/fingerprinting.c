/* fingerprinting.c — generate_fingerprint / generate_fingerprint_from_samples.
 *
 * Drop-in for the reference's fingerprinting.c:10-109.  The stage sequence of the reference
 * (read + down-mix, resample, normalise, spectral images, Haar, raw fingerprints, MinHash) runs as
 * one GPU batch through mnemo_index_pcm_batch; this file keeps the surface: error codes, ownership
 * and the progress lines on stderr (wav.c:346,379,390; fingerprinting.c:19-27,41,58-62,74).
 */
#include <pthread.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <sys/types.h>
#include <time.h>
#include <unistd.h>

#include "fingerprinting.h"
#include "host_internal.h"

/* MNEMO_TRACE=1: wall-clock stage times on stderr (diagnostics only) */
static double now_ms(void) {
    struct timespec ts;
    clock_gettime(CLOCK_MONOTONIC, &ts);
    return ts.tv_sec * 1e3 + ts.tv_nsec / 1e6;
}
static int trace_on(void) {
    static int on = -1;
    if (on < 0) on = getenv("MNEMO_TRACE") != NULL;
    return on;
}
#define TRACE(what, t0) do { if (trace_on()) fprintf(stderr, "[mnemo] %-28s %8.1f ms\n", what, now_ms() - (t0)); } while (0)

/* MNEMO_READERS=n caps the reader threads of the streaming ingest (default: one per core, at most 32) */
static int const_env_readers(void) {
    static int n = -1;
    if (n < 0) {
        const char* e = getenv("MNEMO_READERS");
        n = e ? atoi(e) : 0;
    }
    return n;
}

static pthread_once_t g_once = PTHREAD_ONCE_INIT;
static mnemo_ctx* g_ctx = NULL;

static void make_context(void) {
    const char* dev = getenv("MNEMO_DEVICE");
    if (mnemo_create(dev ? atoi(dev) : 0, NULL, &g_ctx) != 0) g_ctx = NULL;
}

mnemo_ctx* mnemo_host_context(void) {
    pthread_once(&g_once, make_context);
    return g_ctx;
}

/* Creating the CUDA context takes a second or more in a fresh process; it does not depend on the input, so it is
 * started on a helper thread while the caller parses and reads the WAV file. */
static void* warm_context(void* arg) {
    (void)arg;
    mnemo_host_context();
    return NULL;
}
void mnemo_host_context_prefetch(void) {
    static int started = 0;
    if (__sync_lock_test_and_set(&started, 1)) return;
    pthread_t th;
    if (pthread_create(&th, NULL, warm_context, NULL) == 0) pthread_detach(th);
}

void free_signatures(struct signatures* s) {
    if (!s) return;
    free(s->signatures);
    free(s);
}

static struct signatures* new_signatures(uint64_t capacity) {
    struct signatures* s = (struct signatures*)malloc(sizeof *s);
    if (!s) return NULL;
    s->n_signatures = 0;
    s->signatures = (struct signature*)malloc((capacity ? capacity : 1) * sizeof(struct signature));
    if (!s->signatures) {
        free(s);
        return NULL;
    }
    return s;
}

/* Pinned staging for the streaming ingest: a ring of buffers per process, reused for every file. */
#define PIECE_BYTES (4u << 20)
#define RING 64                      /* 256 MB of pinned memory: half in flight to the GPU, half being refilled */
#define MAX_READERS 32              /* one per core at most (24-core boxes: 16 readers left PCIe a third idle) */
static void* g_ring[RING];
static pthread_mutex_t g_stage_lock = PTHREAD_MUTEX_INITIALIZER;   /* one file at a time through the ring */

struct ingest {
    int fd;
    off_t data_off;
    uint64_t total;
    uint32_t n_pieces;
    pthread_mutex_t mu;
    pthread_cond_t cv;
    uint32_t next_piece;    /* next piece a reader may claim */
    uint32_t free_upto;     /* pieces below this index have a free ring slot (= pieces uploaded + RING) */
    uint8_t* state;         /* per piece: 0 pending, 1 in its ring slot, 2 could not be read */
    int stop;
};

static void* ingest_reader(void* arg) {
    struct ingest* g = (struct ingest*)arg;
    for (;;) {
        pthread_mutex_lock(&g->mu);
        const uint32_t i = g->next_piece;
        if (i >= g->n_pieces || g->stop) {
            pthread_mutex_unlock(&g->mu);
            return NULL;
        }
        g->next_piece = i + 1;
        while (i >= g->free_upto && !g->stop) pthread_cond_wait(&g->cv, &g->mu);
        const int stop = g->stop;
        pthread_mutex_unlock(&g->mu);
        if (stop) return NULL;
        const uint64_t off = (uint64_t)i * PIECE_BYTES;
        const size_t want = (size_t)((g->total - off < PIECE_BYTES) ? (g->total - off) : PIECE_BYTES);
        char* dst = (char*)g_ring[i % RING];
        size_t got = 0;
        while (got < want) {
            const ssize_t r = pread(g->fd, dst + got, want - got, g->data_off + (off_t)(off + got));
            if (r <= 0) break;   /* truncated data chunk or I/O error */
            got += (size_t)r;
        }
        pthread_mutex_lock(&g->mu);
        g->state[i] = got == want ? 1 : 2;
        pthread_cond_broadcast(&g->cv);
        pthread_mutex_unlock(&g->mu);
    }
}

/* The data chunk goes to the GPU in 8 MB pieces through a ring of pinned buffers: up to eight reader threads pread
 * the pieces (the page cache delivers ~6 GB/s per thread, PCIe takes 55), the calling thread hands each piece to an
 * asynchronous copy as soon as it is complete and releases a ring slot once its copy has left it.  This replaces
 * read_samples' loop of one fread per sample block (wav.c:358-375). */
static int ingest_streaming(mnemo_ctx* ctx, FILE* f, const struct wav_data* w, struct signatures* s, uint32_t n_images,
                            uint32_t* n_out) {
    mnemo_track tr;
    tr.pcm = NULL;
    tr.n_frames = w->n_frames;
    tr.channels = w->channels;
    tr.reserved = 0;
    mnemo_batch* b = NULL;
    double tc = now_ms();
    int rc = mnemo_batch_create(ctx, &tr, 1, &b);
    TRACE("batch create (device buffers)", tc);
    if (rc != 0) return rc;
    struct ingest g;
    memset(&g, 0, sizeof g);
    g.fd = fileno(f);
    g.data_off = ftello(f);
    g.total = w->n_frames * (uint64_t)w->channels * 2u;
    g.n_pieces = (uint32_t)((g.total + PIECE_BYTES - 1) / PIECE_BYTES);
    g.free_upto = RING;
    g.state = (uint8_t*)calloc(g.n_pieces ? g.n_pieces : 1, 1);
    uint64_t* ticket = (uint64_t*)malloc((g.n_pieces ? g.n_pieces : 1) * sizeof(uint64_t));
    if (!g.state || !ticket || g.data_off < 0) {
        free(g.state);
        free(ticket);
        mnemo_batch_destroy(b);
        return g.data_off < 0 ? CANNOT_READ_FILE : MEMORY_ERROR;
    }
    pthread_mutex_init(&g.mu, NULL);
    pthread_cond_init(&g.cv, NULL);
    pthread_mutex_lock(&g_stage_lock);
    tc = now_ms();
    for (int i = 0; i < RING && rc == 0; i++)
        if (!g_ring[i]) rc = mnemo_host_alloc(ctx, &g_ring[i], PIECE_BYTES);
    TRACE("pinned ring", tc);
    double t0 = now_ms();
    pthread_t th[MAX_READERS];
    int n_readers = 0;
    if (rc == 0) {
        long cores = sysconf(_SC_NPROCESSORS_ONLN);
        int want = (int)(g.n_pieces < MAX_READERS ? g.n_pieces : MAX_READERS);
        if (cores > 0 && want > cores) want = (int)cores;
        if (const_env_readers() > 0 && want > const_env_readers()) want = const_env_readers();
        for (; n_readers < want; n_readers++)
            if (pthread_create(&th[n_readers], NULL, ingest_reader, &g) != 0) break;
        if (n_readers == 0 && g.n_pieces) rc = MEMORY_ERROR;
    }
    for (uint32_t i = 0; rc == 0 && i < g.n_pieces; i++) {
        pthread_mutex_lock(&g.mu);
        while (g.state[i] == 0) pthread_cond_wait(&g.cv, &g.mu);
        const int st = g.state[i];
        pthread_mutex_unlock(&g.mu);
        if (st != 1) {   /* read_samples: DECODING_ERROR on a short data chunk (wav.c:381-386) */
            rc = DECODING_ERROR;
            break;
        }
        const uint64_t off = (uint64_t)i * PIECE_BYTES;
        const uint64_t want = (g.total - off < PIECE_BYTES) ? (g.total - off) : PIECE_BYTES;
        rc = mnemo_batch_upload_part_ticket(b, 0, off, g_ring[i % RING], want, &ticket[i]);
        if (rc == 0 && i >= RING / 2) {   /* half of the ring in flight, half being refilled */
            rc = mnemo_batch_upload_wait(b, ticket[i - RING / 2]);
            pthread_mutex_lock(&g.mu);
            g.free_upto = i - RING / 2 + 1 + RING;
            pthread_cond_broadcast(&g.cv);
            pthread_mutex_unlock(&g.mu);
        }
    }
    pthread_mutex_lock(&g.mu);
    g.stop = 1;
    pthread_cond_broadcast(&g.cv);
    pthread_mutex_unlock(&g.mu);
    for (int i = 0; i < n_readers; i++) pthread_join(th[i], NULL);
    if (rc == 0) rc = mnemo_batch_sync(b);
    else mnemo_batch_sync(b);   /* nothing may still read the ring when the next file takes it */
    pthread_mutex_unlock(&g_stage_lock);
    pthread_cond_destroy(&g.cv);
    pthread_mutex_destroy(&g.mu);
    free(g.state);
    free(ticket);
    TRACE("read + upload (streamed)", t0);
    t0 = now_ms();
    int32_t status = 0;
    if (rc == 0) rc = mnemo_batch_run(b);
    if (rc == 0) rc = mnemo_batch_download(b, (uint8_t*)s->signatures, n_images, n_out, &status);
    TRACE("kernels + download", t0);
    t0 = now_ms();
    mnemo_batch_destroy(b);
    TRACE("batch destroy", t0);
    if (rc == 0) rc = status;
    return rc;
}

int generate_fingerprint(const char* wav, struct signatures** fingerprint,
                         char** artist, char** track_title, char** album_title) {
    struct wav_data w;
    FILE* f = NULL;
    unsigned int align = 0;
    double t0 = now_ms();
    mnemo_host_context_prefetch();
    int rc = wav_open(wav, &w, &f, &align);
    if (rc != SUCCESS) return rc;
    if (w.artist) fprintf(stderr, "Artist: %s\n", w.artist);
    if (w.track_title) fprintf(stderr, "Track title: %s\n", w.track_title);
    if (w.album_title) fprintf(stderr, "Album title: %s\n", w.album_title);
    /* the caller takes the metadata strings (or they are dropped with the reader, as upstream) */
    if (artist) { *artist = w.artist; w.artist = NULL; }
    if (track_title) { *track_title = w.track_title; w.track_title = NULL; }
    if (album_title) { *album_title = w.album_title; w.album_title = NULL; }

    fprintf(stderr, "Reading 44100Hz samples...\n");
    fprintf(stderr, "Resampling to 5512Hz...\n");
    fprintf(stderr, "Normalizing samples...\n");
    const uint32_t n5512 = mnemo_samples_5512(w.n_frames);
    const uint32_t n_images = mnemo_images(n5512);
    const int streamed = (align == 2u * w.channels) && n_images != 0;
    if (!streamed) {
        /* padded sample blocks, or a file too small to fingerprint: the whole-file reader (it also reports a
         * truncated data chunk before the size check, as read_samples does) */
        fclose(f);
        f = NULL;
        char *a2 = w.artist, *t2 = w.track_title, *l2 = w.album_title;
        w.artist = w.track_title = w.album_title = NULL;
        wav_release(&w);
        rc = wav_load(wav, &w);
        /* metadata was already handed out (or kept) above: drop the second copy */
        free(w.artist); free(w.track_title); free(w.album_title);
        w.artist = a2; w.track_title = t2; w.album_title = l2;
        if (rc != SUCCESS) {
            wav_release(&w);
            return rc;
        }
    }
    TRACE("wav header / load", t0);
    fprintf(stderr, "%d 5512Hz mono samples\n", (int)n5512);
    if (n_images == 0) {   /* fingerprinting.c:42, spectralimages.c:128 */
        if (f) fclose(f);
        wav_release(&w);
        return FILE_TOO_SMALL;
    }
    t0 = now_ms();
    mnemo_ctx* ctx = mnemo_host_context();
    TRACE("context (wait)", t0);
    if (!ctx) {
        fprintf(stderr, "mnemophonix: no usable B200 GPU (there is no CPU implementation)\n");
        if (f) fclose(f);
        wav_release(&w);
        return DECODING_ERROR;
    }
    struct signatures* s = new_signatures(n_images);
    if (!s) {
        if (f) fclose(f);
        wav_release(&w);
        return MEMORY_ERROR;
    }
    fprintf(stderr, "Got %d spectral images\n", (int)n_images);
    fprintf(stderr, "Applying Haar transform to spectral images\n");
    fprintf(stderr, "Building raw fingerprints\n");
    uint32_t n = 0;
    if (streamed) {
        rc = ingest_streaming(ctx, f, &w, s, n_images, &n);
        fclose(f);
    } else {
        mnemo_track tr;
        tr.pcm = w.pcm;
        tr.n_frames = w.n_frames;
        tr.channels = w.channels;
        tr.reserved = 0;
        int32_t status = 0;
        t0 = now_ms();
        rc = mnemo_index_pcm_batch(ctx, &tr, 1, (uint8_t*)s->signatures, n_images, &n, &status);
        TRACE("mnemo_index_pcm_batch", t0);
        if (rc == 0) rc = status;
    }
    wav_release(&w);
    if (rc != SUCCESS) {
        free_signatures(s);
        return rc;
    }
    s->n_signatures = n;
    fprintf(stderr, "Generated %d signatures\n", s->n_signatures);
    *fingerprint = s;
    return SUCCESS;
}

int generate_fingerprint_from_samples(float* samples, unsigned int size, struct signatures** fingerprint) {
    const uint32_t n_images = mnemo_images(size);
    if (n_images == 0) return FILE_TOO_SMALL;   /* fingerprinting.c:82, spectralimages.c:128 */
    mnemo_ctx* ctx = mnemo_host_context();
    if (!ctx) return DECODING_ERROR;
    struct signatures* s = new_signatures(n_images);
    if (!s) return MEMORY_ERROR;
    uint32_t n = 0;
    int rc = mnemo_index_samples(ctx, samples, size, (uint8_t*)s->signatures, n_images, &n);
    if (rc != SUCCESS) {
        free_signatures(s);
        return rc;
    }
    s->n_signatures = n;
    *fingerprint = s;
    return SUCCESS;
}
