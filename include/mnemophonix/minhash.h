/* minhash.h — signature types of the mnemophonix C API (layout identical to the reference's
 * minhash.h:9-29: 100-byte signature, {unsigned n; pointer} container). */
#ifndef MNEMOPHONIX_MINHASH_H
#define MNEMOPHONIX_MINHASH_H
#include <stdint.h>

#define SIGNATURE_LENGTH 100

struct signature {
    uint8_t minhash[SIGNATURE_LENGTH]; /* 0..254 = first set bit along permutation i, 255 = none */
};

struct signatures {
    unsigned int n_signatures;
    struct signature* signatures;
};

/* Releases what generate_fingerprint* / read_index allocated (reference: minhash.c:57-60). */
void free_signatures(struct signatures* signatures);

#endif
