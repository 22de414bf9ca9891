/* lsh_search.c — create_hash_tables / get_matches / search (drop-ins for lsh.c:55-112 and
 * search.c:110-193) over the device-resident tables of libmnemo_cuda.so.
 *
 * struct lsh keeps the reference's public layout (size, buckets[25]); the device handle sits in a
 * private tail allocated behind it.  Concurrent search() calls on one lsh are serialised on the
 * context's stream by a mutex (the reference's search is re-entrant; so is this one).
 */
#include <pthread.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "host_internal.h"
#include "lsh.h"
#include "search.h"

struct lsh_private {
    struct lsh pub;     /* must be first */
    mnemo_db* db;
    pthread_mutex_t lock;
};

struct signature_list* new_signature_list(unsigned int entry_index, unsigned int signature_index,
                                          struct signature_list* next) {
    struct signature_list* n = (struct signature_list*)malloc(sizeof *n);
    if (!n) return NULL;
    n->entry_index = entry_index;
    n->signature_index = signature_index;
    n->next = next;
    return n;
}

void free_signature_list(struct signature_list* list) {
    while (list) {
        struct signature_list* next = list->next;
        free(list);
        list = next;
    }
}

struct lsh* create_hash_tables(struct index* database) {
    mnemo_ctx* ctx = mnemo_host_context();
    if (!ctx || !database) return NULL;
    struct lsh_private* p = (struct lsh_private*)calloc(1, sizeof *p);
    if (!p) return NULL;
    const unsigned int n = database->n_entries;
    uint32_t* start = (uint32_t*)malloc(((size_t)n + 1) * sizeof(uint32_t));
    if (!start) {
        free(p);
        return NULL;
    }
    start[0] = 0;
    for (unsigned int i = 0; i < n; i++) start[i + 1] = start[i] + database->entries[i]->signatures->n_signatures;
    const uint32_t total = start[n];
    uint8_t* all = (uint8_t*)malloc((size_t)(total ? total : 1) * SIGNATURE_LENGTH);
    if (!all) {
        free(start);
        free(p);
        return NULL;
    }
    for (unsigned int i = 0; i < n; i++)
        memcpy(all + (size_t)start[i] * SIGNATURE_LENGTH, database->entries[i]->signatures->signatures,
               (size_t)(start[i + 1] - start[i]) * SIGNATURE_LENGTH);
    int rc = mnemo_db_create(ctx, all, start, n, 0, 0, &p->db);
    free(all);
    free(start);
    if (rc != 0) {
        free(p);
        return NULL;
    }
    p->pub.size = mnemo_db_table_size(p->db);   /* lsh.c:61 */
    pthread_mutex_init(&p->lock, NULL);
    return &p->pub;
}

void free_hash_tables(struct lsh* tables) {
    if (!tables) return;
    struct lsh_private* p = (struct lsh_private*)tables;
    mnemo_db_destroy(p->db);
    pthread_mutex_destroy(&p->lock);
    free(p);
}

int get_matches(struct lsh* tables, uint8_t* hash, struct signature_list** list) {
    struct lsh_private* p = (struct lsh_private*)tables;
    *list = NULL;
    uint64_t cap = 1u << 14;
    uint32_t* pairs = NULL;
    int64_t n;
    for (;;) {
        pairs = (uint32_t*)malloc(cap * 2 * sizeof(uint32_t));
        if (!pairs) return MEMORY_ERROR;
        pthread_mutex_lock(&p->lock);
        n = mnemo_db_get_matches(p->db, hash, pairs, cap);
        pthread_mutex_unlock(&p->lock);
        if (n < 0) {
            free(pairs);
            return MEMORY_ERROR;
        }
        if ((uint64_t)n <= cap) break;
        free(pairs);
        cap = (uint64_t)n;
    }
    /* build the list back to front so that it reads in the returned order */
    for (int64_t i = n - 1; i >= 0; i--) {
        struct signature_list* node = new_signature_list(pairs[2 * i], pairs[2 * i + 1], *list);
        if (!node) {
            free_signature_list(*list);
            *list = NULL;
            free(pairs);
            return MEMORY_ERROR;
        }
        *list = node;
    }
    free(pairs);
    return (int)n;
}

/* Text of the last CUDA-side failure behind this library (empty when there was none): search() and
 * create_hash_tables() keep the reference's return conventions, this tells a GPU fault from a miss. */
const char* mnemo_host_last_error(void) {
    mnemo_ctx* ctx = mnemo_host_context();
    return ctx ? mnemo_last_error(ctx) : "no sm_100 GPU available (there is no CPU path)";
}

int search(struct signatures* sample, struct index* database, struct lsh* lsh, int verbose) {
    struct lsh_private* p = (struct lsh_private*)lsh;
    if (!p || !sample || !database) return MEMORY_ERROR;
    uint32_t qstart[2] = {0, sample->n_signatures};
    mnemo_match m = {NO_MATCH_FOUND, 0.0f, 0};
    mnemo_sel_rec top[10];
    uint32_t n_top = 0;
    pthread_mutex_lock(&p->lock);
    int rc = verbose ? mnemo_search_batch_top(p->db, (const uint8_t*)sample->signatures, qstart, 1, &m, top, &n_top)
                     : mnemo_search_batch(p->db, (const uint8_t*)sample->signatures, qstart, 1, &m, NULL);
    pthread_mutex_unlock(&p->lock);
    if (rc != 0) return MEMORY_ERROR;
    if (verbose) {
        /* search.c:175-178: the first ten records of the sorted scores[] array.  Matched entries come first, in the
         * order of the reference's merge sort (n_top of them, from the GPU); the entries without matches follow in
         * index order (they compare equal, and glibc's merge sort is stable). */
        unsigned int printed = 0;
        for (; printed < n_top && printed < 10; printed++)
            printf("average_score = %f, n_matches = %d (%s)\n", top[printed].avg, top[printed].n_matches,
                   database->entries[top[printed].entry]->filename);
        for (unsigned int e = 0; e < database->n_entries && printed < 10; e++) {
            int matched = 0;
            for (unsigned int k = 0; k < n_top; k++) matched |= top[k].entry == (int32_t)e;
            if (matched) continue;
            printf("average_score = %f, n_matches = %d (%s)\n", 0.0f, 0, database->entries[e]->filename);
            printed++;
        }
        if (m.entry >= 0)
            printf("\n*** match = %f %s ***\n", m.n_matches ? m.score / (float)m.n_matches : 0.0f,
                   database->entries[m.entry]->filename);
        printf("-----------------------------\n");
    }
    return m.entry;
}
