/* wavfile.c — RIFF/WAVE container parsing and bulk PCM ingest.
 *
 * Behaviour follows the reference's wav.c (new_wav_reader wav.c:258-299, read_format_chunk
 * :118-146, skip_optional_chunks :158-178, read_info :188-255, read_samples :345-394):
 *   - file cannot be opened                          -> CANNOT_READ_FILE
 *   - no "RIFF"<size>"WAVE" header                   -> NOT_A_WAVE_FILE
 *   - next chunk is not "fmt " or is shorter than 16 -> DECODING_ERROR
 *   - not PCM / fmt size != 16 / not 44100 Hz / not 16 bit -> UNSUPPORTED_WAVE_FORMAT
 *   - chunks are skipped until "data"; a trailing LIST/INFO chunk yields IART / INAM / IPRD
 *   - the sample count is data bytes / block align; a short data chunk -> DECODING_ERROR
 * What differs is the mechanics: the data chunk is read with ONE fread into memory (the reference
 * issues one fread per sample block, a third of its indexing time), because the down-mix itself
 * happens on the GPU.
 */
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "errors.h"
#include "host_internal.h"

static int rd(FILE* f, void* dst, size_t n) { return fread(dst, 1, n, f) == n; }

static int rd_u32(FILE* f, uint32_t* v) {
    uint8_t b[4];
    if (!rd(f, b, 4)) return 0;
    *v = (uint32_t)b[0] | ((uint32_t)b[1] << 8) | ((uint32_t)b[2] << 16) | ((uint32_t)b[3] << 24);
    return 1;
}

static int rd_u16(FILE* f, uint16_t* v) {
    uint8_t b[2];
    if (!rd(f, b, 2)) return 0;
    *v = (uint16_t)(b[0] | (b[1] << 8));
    return 1;
}

static int is_tag(const uint8_t* t, const char* name) { return memcmp(t, name, 4) == 0; }

void wav_release(struct wav_data* w) {
    free(w->pcm);
    free(w->artist);
    free(w->track_title);
    free(w->album_title);
    memset(w, 0, sizeof *w);
}

/* LIST/INFO after the data chunk (wav.c:188-255). */
static int parse_info(FILE* f, uint32_t file_size, uint32_t after_data, struct wav_data* w) {
    uint8_t tag[4];
    uint32_t size;
    if (after_data == file_size) return SUCCESS;
    if (fseek(f, (long)after_data, SEEK_SET) != 0) return DECODING_ERROR;
    if (!rd(f, tag, 4) || !rd_u32(f, &size)) return DECODING_ERROR;
    if (!is_tag(tag, "LIST")) return SUCCESS;
    if (!rd(f, tag, 4)) return DECODING_ERROR;
    if (!is_tag(tag, "INFO")) return SUCCESS;
    uint32_t left = size - 4;
    while (left > 0) {
        uint32_t n;
        if (!rd(f, tag, 4) || !rd_u32(f, &n)) return DECODING_ERROR;
        char** slot = is_tag(tag, "IART") ? &w->artist : is_tag(tag, "INAM") ? &w->track_title
                      : is_tag(tag, "IPRD") ? &w->album_title : NULL;
        if (slot) {
            char* s = (char*)malloc((size_t)n + 1);
            if (!s) return MEMORY_ERROR;
            if (!rd(f, s, n)) {
                free(s);
                return DECODING_ERROR;
            }
            s[n] = '\0';
            free(*slot);
            *slot = s;
        } else if (fseek(f, (long)n, SEEK_CUR) != 0) {
            return DECODING_ERROR;
        }
        left -= 8 + n;   /* unsigned wrap-around ends the loop like the reference's does not: guard */
        if (left > size) break;
    }
    return SUCCESS;
}

int wav_open(const char* path, struct wav_data* w, FILE** f_out, unsigned int* block_align) {
    memset(w, 0, sizeof *w);
    FILE* f = fopen(path, "rb");
    if (!f) return CANNOT_READ_FILE;
    int rc = SUCCESS;
    uint8_t tag[4], form[4];
    uint32_t riff_size = 0;
    if (!rd(f, tag, 4) || !is_tag(tag, "RIFF") || !rd_u32(f, &riff_size) || !rd(f, form, 4) || !is_tag(form, "WAVE")) {
        fclose(f);
        return NOT_A_WAVE_FILE;
    }
    const uint32_t file_size = riff_size + 8;
    uint32_t fmt_size = 0, rate = 0, avg = 0;
    uint16_t format = 0, channels = 0, align = 0, bits = 0;
    if (!rd(f, tag, 4) || !is_tag(tag, "fmt ") || !rd_u32(f, &fmt_size) || fmt_size < 16 || !rd_u16(f, &format) ||
        !rd_u16(f, &channels) || !rd_u32(f, &rate) || !rd_u32(f, &avg) || !rd_u16(f, &align) || !rd_u16(f, &bits)) {
        fclose(f);
        return DECODING_ERROR;
    }
    (void)avg;
    if ((int16_t)format != 1 || fmt_size != 16 || rate != 44100 || bits != 16) {
        fclose(f);
        return UNSUPPORTED_WAVE_FORMAT;
    }
    uint32_t data_size = 0;
    for (;;) {   /* skip chunks up to "data" */
        if (!rd(f, tag, 4) || !rd_u32(f, &data_size)) {
            fclose(f);
            return DECODING_ERROR;
        }
        if (is_tag(tag, "data")) break;
        if (fseek(f, (long)data_size, SEEK_CUR) != 0) {
            fclose(f);
            return DECODING_ERROR;
        }
    }
    const long data_pos = ftell(f);
    rc = parse_info(f, file_size, (uint32_t)data_pos + data_size, w);
    if (rc != SUCCESS) goto fail;
    if (align == 0 || channels == 0 || (unsigned)align < 2u * channels) {   /* the reference would crash or read garbage */
        rc = DECODING_ERROR;
        goto fail;
    }
    w->channels = channels;
    w->n_frames = data_size / align;   /* wav.c:347 */
    if (fseek(f, data_pos, SEEK_SET) != 0) {
        rc = DECODING_ERROR;
        goto fail;
    }
    *f_out = f;
    *block_align = align;
    return SUCCESS;
fail:
    fclose(f);
    wav_release(w);
    return rc;
}

int wav_load(const char* path, struct wav_data* w) {
    FILE* f = NULL;
    unsigned int align = 0;
    int rc = wav_open(path, w, &f, &align);
    if (rc != SUCCESS) return rc;
    const unsigned int channels = w->channels;
    {
        const size_t bytes = (size_t)w->n_frames * align;
        uint8_t* raw = (uint8_t*)malloc(bytes ? bytes : 1);
        if (!raw) {
            rc = MEMORY_ERROR;
            goto fail;
        }
        if (!rd(f, raw, bytes)) {   /* truncated data chunk: read_samples returns DECODING_ERROR */
            free(raw);
            rc = DECODING_ERROR;
            goto fail;
        }
        if ((unsigned)align == 2u * channels) {
            w->pcm = (int16_t*)raw;   /* little-endian host: the file layout is the kernel's layout */
        } else {                      /* padded blocks: keep the first `channels` samples of each */
            int16_t* packed = (int16_t*)malloc((size_t)w->n_frames * channels * 2 + 2);
            if (!packed) {
                free(raw);
                rc = MEMORY_ERROR;
                goto fail;
            }
            for (uint64_t i = 0; i < w->n_frames; i++)
                memcpy(packed + i * channels, raw + i * align, (size_t)channels * 2);
            free(raw);
            w->pcm = packed;
        }
    }
    fclose(f);
    return SUCCESS;
fail:
    fclose(f);
    wav_release(w);
    return rc;
}
