/* fingerprintio.h — text database format, drop-in for the reference's fingerprintio.h:11-68.
 * An entry is 5 header lines (name, artist, title, album, signature count) followed by one line of
 * 200 lower-case hex characters per signature; a database is a concatenation of entries. */
#ifndef MNEMOPHONIX_FINGERPRINTIO_H
#define MNEMOPHONIX_FINGERPRINTIO_H
#include <stdio.h>
#include "errors.h"
#include "minhash.h"

struct index_entry {
    char* filename;
    char* artist;
    char* track_title;
    char* album_title;
    struct signatures* signatures;
};

struct index {
    unsigned int n_entries;
    struct index_entry** entries;
};

void save(FILE* f, struct signatures* fingerprint, const char* wavname,
          const char* artist, const char* track_title, const char* album_title);
int read_index(const char* filename, struct index** index);
void free_index(struct index* index);

#endif
