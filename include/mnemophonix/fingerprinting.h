/* fingerprinting.h — drop-in for the reference's fingerprinting.h:26-41.  Same prototypes, same
 * ownership (callee mallocs *fingerprint and the metadata strings, caller frees), same return
 * codes.  The body runs on the GPU through libmnemo_cuda.so. */
#ifndef MNEMOPHONIX_FINGERPRINTING_H
#define MNEMOPHONIX_FINGERPRINTING_H
#include "errors.h"
#include "minhash.h"

/* wav: path of a 44.1 kHz 16-bit PCM WAVE file (any channel count).  artist / track_title /
 * album_title may be NULL; otherwise they receive malloc'd strings or NULL (LIST/INFO chunk). */
int generate_fingerprint(const char* wav, struct signatures** fingerprint,
                         char** artist, char** track_title, char** album_title);

/* samples: mono, 5512 Hz, already normalised floats. */
int generate_fingerprint_from_samples(float* samples, unsigned int size, struct signatures** fingerprint);

#endif
