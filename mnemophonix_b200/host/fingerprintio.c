/* fingerprintio.c — the plain-text signature database (drop-in for fingerprintio.c:10-227).
 *
 * Format (fingerprintio.h:31-41 upstream): per entry the lines name, artist, title, album,
 * signature count, then one line of exactly 200 hex characters per signature.  Lines are limited to
 * 1023 characters; end of file at an entry boundary ends the database; anything else malformed is a
 * DECODING_ERROR.  The reference parses every byte with sscanf("%2x") (two thirds of its search
 * start-up time); here a 256-entry nibble table does it, and a well-formed file is parsed by several threads:
 * signature lines have a fixed length, so one sequential walk over the headers finds every entry (skipping its
 * signature block), and the blocks are decoded in parallel from the mapped file.  Anything unexpected makes that
 * fast path give up and the sequential parser, which reproduces the reference's error codes, reads the file again.
 */
#include <fcntl.h>
#include <pthread.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <sys/mman.h>
#include <sys/stat.h>
#include <unistd.h>

#include "fingerprintio.h"

#define LINE_MAX_LEN 1024

void save(FILE* f, struct signatures* fingerprint, const char* wavname,
          const char* artist, const char* track_title, const char* album_title) {
    static const char hex[] = "0123456789abcdef";
    fprintf(f, "%s\n%s\n%s\n%s\n%d\n", wavname, artist ? artist : "", track_title ? track_title : "",
            album_title ? album_title : "", fingerprint->n_signatures);
    char line[2 * SIGNATURE_LENGTH + 1];
    line[2 * SIGNATURE_LENGTH] = '\n';
    for (unsigned int i = 0; i < fingerprint->n_signatures; i++) {
        const uint8_t* m = fingerprint->signatures[i].minhash;
        for (int j = 0; j < SIGNATURE_LENGTH; j++) {
            line[2 * j] = hex[m[j] >> 4];
            line[2 * j + 1] = hex[m[j] & 15];
        }
        fwrite(line, 1, sizeof line, f);
    }
}

/* Reads one line without its '\n'.  SUCCESS, CANNOT_READ_FILE at end of file, DECODING_ERROR if the
 * line is too long or unterminated. */
static int next_line(FILE* f, char* buf, size_t* len) {
    if (!fgets(buf, LINE_MAX_LEN, f)) return CANNOT_READ_FILE;
    size_t n = strlen(buf);
    if (n == 0 || buf[n - 1] != '\n') return DECODING_ERROR;
    buf[--n] = '\0';
    if (len) *len = n;
    return SUCCESS;
}

static void free_entry(struct index_entry* e) {
    if (!e) return;
    free(e->filename);
    free(e->artist);
    free(e->track_title);
    free(e->album_title);
    if (e->signatures) free_signatures(e->signatures);
    free(e);
}

static int8_t g_nibble[256];
static pthread_once_t g_nibble_once = PTHREAD_ONCE_INIT;   /* read_index is re-entrant like the reference's */

static void init_nibbles(void) {
    memset(g_nibble, -1, sizeof g_nibble);
    for (int c = '0'; c <= '9'; c++) g_nibble[c] = (int8_t)(c - '0');
    for (int c = 'a'; c <= 'f'; c++) g_nibble[c] = (int8_t)(c - 'a' + 10);
    for (int c = 'A'; c <= 'F'; c++) g_nibble[c] = (int8_t)(c - 'A' + 10);
}

static int read_one_entry(FILE* f, struct index_entry** out) {
    char buf[LINE_MAX_LEN];
    size_t len;
    int rc = next_line(f, buf, NULL);
    if (rc == CANNOT_READ_FILE) return CANNOT_READ_FILE;   /* clean end of the database */
    /* fingerprintio.c:78-80 upstream checks only for the end of the file here: an over-long or unterminated first line
     * is kept as it was read (buf is untouched on DECODING_ERROR) and parsing goes on with what follows */
    struct index_entry* e = (struct index_entry*)calloc(1, sizeof *e);
    if (!e) return MEMORY_ERROR;
    char** fields[4] = {&e->filename, &e->artist, &e->track_title, &e->album_title};
    for (int k = 0; k < 4; k++) {
        if (k > 0 && (rc = next_line(f, buf, NULL)) != SUCCESS) goto bad;
        if (!(*fields[k] = strdup(buf))) {
            rc = MEMORY_ERROR;
            goto bad;
        }
    }
    if ((rc = next_line(f, buf, NULL)) != SUCCESS) goto bad;
    unsigned int n;
    if (sscanf(buf, "%u", &n) != 1) {
        rc = DECODING_ERROR;
        goto bad;
    }
    e->signatures = (struct signatures*)malloc(sizeof(struct signatures));
    if (!e->signatures) {
        rc = MEMORY_ERROR;
        goto bad;
    }
    e->signatures->n_signatures = n;
    e->signatures->signatures = (struct signature*)malloc((n ? n : 1) * sizeof(struct signature));
    if (!e->signatures->signatures) {
        free(e->signatures);
        e->signatures = NULL;
        rc = MEMORY_ERROR;
        goto bad;
    }
    for (unsigned int i = 0; i < n; i++) {
        if ((rc = next_line(f, buf, &len)) != SUCCESS) goto bad;
        if (len != 2 * SIGNATURE_LENGTH) {
            rc = DECODING_ERROR;
            goto bad;
        }
        uint8_t* m = e->signatures->signatures[i].minhash;
        for (int j = 0; j < SIGNATURE_LENGTH; j++) {
            const int hi = g_nibble[(unsigned char)buf[2 * j]], lo = g_nibble[(unsigned char)buf[2 * j + 1]];
            if ((hi | lo) < 0) {
                rc = DECODING_ERROR;
                goto bad;
            }
            m[j] = (uint8_t)((hi << 4) | lo);
        }
    }
    *out = e;
    return SUCCESS;
bad:
    free_entry(e);
    return rc == CANNOT_READ_FILE ? DECODING_ERROR : rc;   /* end of file inside an entry */
}

/* ---- fast path: mapped file, entries located sequentially, signature blocks decoded by threads ---- */
#define SIG_LINE (2 * SIGNATURE_LENGTH + 1)
#define FAST_MAX_THREADS 8

struct located {
    const char* fields[4];
    size_t field_len[4];
    const char* sigs;      /* n lines of SIG_LINE bytes */
    unsigned int n;
};

struct decode_job {
    const struct located* loc;
    struct index_entry** entries;
    unsigned int first, last;   /* entries [first, last) */
    int failed;
};

static char* dup_range(const char* p, size_t n) {
    char* s = (char*)malloc(n + 1);
    if (!s) return NULL;
    memcpy(s, p, n);
    s[n] = '\0';
    return s;
}

static void* decode_entries(void* arg) {
    struct decode_job* job = (struct decode_job*)arg;
    for (unsigned int i = job->first; i < job->last && !job->failed; i++) {
        const struct located* l = &job->loc[i];
        struct index_entry* e = (struct index_entry*)calloc(1, sizeof *e);
        if (!e) {
            job->failed = 1;
            break;
        }
        job->entries[i] = e;
        e->filename = dup_range(l->fields[0], l->field_len[0]);
        e->artist = dup_range(l->fields[1], l->field_len[1]);
        e->track_title = dup_range(l->fields[2], l->field_len[2]);
        e->album_title = dup_range(l->fields[3], l->field_len[3]);
        e->signatures = (struct signatures*)calloc(1, sizeof(struct signatures));   /* (freeable when half built) */
        if (!e->filename || !e->artist || !e->track_title || !e->album_title || !e->signatures) {
            job->failed = 1;
            break;
        }
        e->signatures->n_signatures = l->n;
        e->signatures->signatures = (struct signature*)malloc((l->n ? l->n : 1) * sizeof(struct signature));
        if (!e->signatures->signatures) {
            job->failed = 1;
            break;
        }
        for (unsigned int k = 0; k < l->n && !job->failed; k++) {
            const unsigned char* line = (const unsigned char*)l->sigs + (size_t)k * SIG_LINE;
            uint8_t* m = e->signatures->signatures[k].minhash;
            int bad = line[2 * SIGNATURE_LENGTH] != '\n';
            for (int j = 0; j < SIGNATURE_LENGTH; j++) {
                const int hi = g_nibble[line[2 * j]], lo = g_nibble[line[2 * j + 1]];
                bad |= (hi | lo) < 0;
                m[j] = (uint8_t)((hi << 4) | lo);
            }
            if (bad) job->failed = 1;
        }
    }
    return NULL;
}

/* Returns 1 and fills *index if the file is well formed; 0 = let the sequential parser decide. */
static int read_index_fast(const char* filename, struct index** index) {
    int fd = open(filename, O_RDONLY);
    if (fd < 0) return 0;
    struct stat st;
    if (fstat(fd, &st) != 0 || !S_ISREG(st.st_mode)) {
        close(fd);
        return 0;
    }
    const size_t size = (size_t)st.st_size;
    const char* base = size ? (const char*)mmap(NULL, size, PROT_READ, MAP_PRIVATE, fd, 0) : "";
    close(fd);
    if (base == (const char*)MAP_FAILED) return 0;
    int ok = 1;
    unsigned int n_entries = 0, cap = 64;
    struct located* loc = (struct located*)malloc(cap * sizeof *loc);
    if (!loc) ok = 0;
    size_t pos = 0;
    while (ok && pos < size) {
        if (n_entries == cap) {
            struct located* bigger = (struct located*)realloc(loc, 2 * (size_t)cap * sizeof *loc);
            if (!bigger) {
                ok = 0;
                break;
            }
            loc = bigger;
            cap *= 2;
        }
        struct located* l = &loc[n_entries];
        const char* lines[5];
        size_t lens[5];
        for (int k = 0; k < 5 && ok; k++) {
            const char* p = base + pos;
            const char* nl = (const char*)memchr(p, '\n', size - pos);
            if (!nl || (size_t)(nl - p) > LINE_MAX_LEN - 2 || memchr(p, '\0', (size_t)(nl - p))) {
                ok = 0;   /* unterminated, over-long (fgets would split it) or embedded NUL */
                break;
            }
            lines[k] = p;
            lens[k] = (size_t)(nl - p);
            pos += lens[k] + 1;
        }
        if (!ok) break;
        /* the count: plain decimal digits only (anything sscanf("%u") would treat differently goes the slow way) */
        unsigned long long n = 0;
        if (lens[4] == 0 || lens[4] > 10) ok = 0;
        for (size_t i = 0; ok && i < lens[4]; i++) {
            if (lines[4][i] < '0' || lines[4][i] > '9') ok = 0;
            n = n * 10 + (unsigned long long)(lines[4][i] - '0');
        }
        if (!ok || n > 0xFFFFFFFFull || n * SIG_LINE > size - pos) {
            ok = 0;
            break;
        }
        for (int k = 0; k < 4; k++) {
            l->fields[k] = lines[k];
            l->field_len[k] = lens[k];
        }
        l->sigs = base + pos;
        l->n = (unsigned int)n;
        pos += (size_t)n * SIG_LINE;
        n_entries++;
    }
    struct index* ix = NULL;
    if (ok) {
        ix = (struct index*)malloc(sizeof *ix);
        if (ix) {
            ix->n_entries = n_entries;
            ix->entries = (struct index_entry**)calloc(n_entries ? n_entries : 1, sizeof *ix->entries);
            if (!ix->entries) {
                free(ix);
                ix = NULL;
            }
        }
        if (!ix) ok = 0;
    }
    if (ok) {
        /* contiguous runs of entries with about the same number of signatures per thread */
        unsigned long long total = 0;
        for (unsigned int i = 0; i < n_entries; i++) total += loc[i].n + 4;
        long cores = sysconf(_SC_NPROCESSORS_ONLN);
        int n_threads = (int)(cores < 1 ? 1 : cores > FAST_MAX_THREADS ? FAST_MAX_THREADS : cores);
        if (total < 200000) n_threads = 1;
        struct decode_job jobs[FAST_MAX_THREADS];
        pthread_t th[FAST_MAX_THREADS];
        int started[FAST_MAX_THREADS];
        unsigned int first = 0;
        unsigned long long acc = 0;
        for (int t = 0; t < n_threads; t++) {
            const unsigned long long goal = total * (unsigned long long)(t + 1) / (unsigned long long)n_threads;
            unsigned int last = first;
            while (last < n_entries && (acc < goal || t == n_threads - 1)) acc += loc[last++].n + 4;
            jobs[t].loc = loc;
            jobs[t].entries = ix->entries;
            jobs[t].first = first;
            jobs[t].last = last;
            jobs[t].failed = 0;
            first = last;
            started[t] = 0;
            if (t + 1 < n_threads && jobs[t].first < jobs[t].last)
                started[t] = pthread_create(&th[t], NULL, decode_entries, &jobs[t]) == 0;
            if (!started[t]) decode_entries(&jobs[t]);   /* the last share (or a thread that would not start) runs here */
        }
        for (int t = 0; t < n_threads; t++) {
            if (started[t]) pthread_join(th[t], NULL);
            if (jobs[t].failed) ok = 0;
        }
        if (!ok) {
            free_index(ix);
            ix = NULL;
        }
    }
    free(loc);
    if (size) munmap((void*)base, size);
    if (ok) *index = ix;
    return ok;
}

int read_index(const char* filename, struct index** index) {
    pthread_once(&g_nibble_once, init_nibbles);
    if (read_index_fast(filename, index)) return SUCCESS;
    FILE* f = fopen(filename, "r");
    if (!f) return CANNOT_READ_FILE;
    struct index* ix = (struct index*)malloc(sizeof *ix);
    if (!ix) {
        fclose(f);
        return MEMORY_ERROR;
    }
    unsigned int cap = 16;
    ix->n_entries = 0;
    ix->entries = (struct index_entry**)malloc(cap * sizeof *ix->entries);
    if (!ix->entries) {
        free(ix);
        fclose(f);
        return MEMORY_ERROR;
    }
    for (;;) {
        struct index_entry* e = NULL;
        int rc = read_one_entry(f, &e);
        if (rc == CANNOT_READ_FILE) break;
        if (rc == SUCCESS && ix->n_entries == cap) {
            struct index_entry** bigger = (struct index_entry**)realloc(ix->entries, 2 * (size_t)cap * sizeof *bigger);
            if (!bigger) {
                free_entry(e);
                rc = MEMORY_ERROR;
            } else {
                ix->entries = bigger;
                cap *= 2;
            }
        }
        if (rc != SUCCESS) {
            free_index(ix);
            fclose(f);
            return rc;
        }
        ix->entries[ix->n_entries++] = e;
    }
    fclose(f);
    *index = ix;
    return SUCCESS;
}

void free_index(struct index* index) {
    if (!index) return;
    for (unsigned int i = 0; i < index->n_entries; i++) free_entry(index->entries[i]);
    free(index->entries);
    free(index);
}
