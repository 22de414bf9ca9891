/* host_internal.h — private glue of libmnemophonix.so (host C over the libmnemo_cuda.so C ABI). */
#ifndef MNEMO_HOST_INTERNAL_H
#define MNEMO_HOST_INTERNAL_H
#include <stdint.h>
#include "mnemo_cuda.h"

/* Process-wide GPU context, created on first use (device = $MNEMO_DEVICE, default 0).
 * Returns NULL if no B200 is available: there is no CPU path. */
mnemo_ctx* mnemo_host_context(void);
/* Starts creating it on a helper thread (returns at once); mnemo_host_context() then waits for it. */
void mnemo_host_context_prefetch(void);

/* One parsed WAVE file: PCM frames in memory plus LIST/INFO metadata (malloc'd, may be NULL). */
struct wav_data {
    int16_t* pcm;          /* interleaved, channels per frame */
    uint64_t n_frames;
    unsigned int channels;
    char* artist;
    char* track_title;
    char* album_title;
};
int wav_load(const char* path, struct wav_data* out);
void wav_release(struct wav_data* w);
/* Same parsing and error codes as wav_load, but the data chunk is left in the file: *f_out is positioned at its
 * first byte, *block_align is the file's frame stride; out->pcm stays NULL.  The caller closes the file. */
#include <stdio.h>
int wav_open(const char* path, struct wav_data* out, FILE** f_out, unsigned int* block_align);

#endif
