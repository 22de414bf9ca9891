/* search.h — drop-in for the reference's search.h:22. */
#ifndef MNEMOPHONIX_SEARCH_H
#define MNEMOPHONIX_SEARCH_H
#include "errors.h"
#include "fingerprintio.h"
#include "lsh.h"
#include "minhash.h"

/* Returns the index of the matching database entry, NO_MATCH_FOUND or MEMORY_ERROR. */
int search(struct signatures* sample, struct index* database, struct lsh* lsh, int verbose);

#endif
