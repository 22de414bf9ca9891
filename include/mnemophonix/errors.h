/* errors.h — return codes of the mnemophonix C API (values identical to the reference's errors.h:5-29). */
#ifndef MNEMOPHONIX_ERRORS_H
#define MNEMOPHONIX_ERRORS_H

#define SUCCESS 0
#define MEMORY_ERROR (-1)            /* allocation failed (host or device) */
#define CANNOT_READ_FILE (-2)        /* file cannot be opened */
#define DECODING_ERROR (-3)          /* input cannot be decoded as expected (also: CUDA failure) */
#define UNSUPPORTED_WAVE_FORMAT (-4) /* a WAVE file, but not PCM / 44100 Hz / 16 bit */
#define FILE_TOO_SMALL (-5)          /* fewer than 2048 samples or 128 frames at 5512 Hz */
#define NOT_A_WAVE_FILE (-6)         /* no RIFF/WAVE header */
#define NO_MATCH_FOUND (-7)          /* search found nothing above the thresholds */

#endif
