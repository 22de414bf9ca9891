/* fingerprintio.c — the plain-text signature database (drop-in for fingerprintio.c:10-227).
 *
 * Format (fingerprintio.h:31-41 upstream): per entry the lines name, artist, title, album,
 * signature count, then one line of exactly 200 hex characters per signature.  Lines are limited to
 * 1023 characters; end of file at an entry boundary ends the database; anything else malformed is a
 * DECODING_ERROR.  The reference parses every byte with sscanf("%2x") (two thirds of its search
 * start-up time); here a 256-entry nibble table does it.
 */
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

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
static int g_nibble_ready = 0;

static void init_nibbles(void) {
    memset(g_nibble, -1, sizeof g_nibble);
    for (int c = '0'; c <= '9'; c++) g_nibble[c] = (int8_t)(c - '0');
    for (int c = 'a'; c <= 'f'; c++) g_nibble[c] = (int8_t)(c - 'a' + 10);
    for (int c = 'A'; c <= 'F'; c++) g_nibble[c] = (int8_t)(c - 'A' + 10);
    g_nibble_ready = 1;
}

static int read_one_entry(FILE* f, struct index_entry** out) {
    char buf[LINE_MAX_LEN];
    size_t len;
    int rc = next_line(f, buf, NULL);
    if (rc == CANNOT_READ_FILE) return CANNOT_READ_FILE;   /* clean end of the database */
    if (rc != SUCCESS) return rc;
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

int read_index(const char* filename, struct index** index) {
    if (!g_nibble_ready) init_nibbles();
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
