/* lsh.h — drop-in for the reference's lsh.h:15-81.  struct layouts are the reference's; the hash
 * tables themselves live on the GPU (a CSR per table), so `buckets` stays NULL and callers go
 * through get_matches(), as every caller of the reference does. */
#ifndef MNEMOPHONIX_LSH_H
#define MNEMOPHONIX_LSH_H
#include <stdint.h>
#include "fingerprintio.h"
#include "minhash.h"

#define BYTES_PER_BUCKET_HASH 4
#define N_BUCKETS (SIGNATURE_LENGTH / BYTES_PER_BUCKET_HASH)

struct signature_list {
    unsigned int entry_index;
    unsigned int signature_index;
    struct signature_list* next;
};

struct lsh {
    unsigned int size;                           /* total signatures / 2 */
    struct signature_list** buckets[N_BUCKETS];  /* unused here: the tables are device-resident */
};

struct lsh* create_hash_tables(struct index* database);
void free_hash_tables(struct lsh* tables);
struct signature_list* new_signature_list(unsigned int entry_index, unsigned int signature_index,
                                          struct signature_list* next);
void free_signature_list(struct signature_list* list);
/* Returns the number of list nodes, or MEMORY_ERROR. */
int get_matches(struct lsh* tables, uint8_t* hash, struct signature_list** list);

#endif
