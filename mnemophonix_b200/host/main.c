/* main.c — the `mnemophonix index|search` command line (surface of the reference's main.c:19-136:
 * same arguments, same stdout/stderr text, same exit status).  Everything heavy happens on the GPU. */
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <signal.h>
#include <sys/socket.h>
#include <sys/time.h>
#include <sys/un.h>
#include <sys/wait.h>
#include <unistd.h>

#include "fingerprinting.h"
#include "fingerprintio.h"
#include "lsh.h"
#include "search.h"

static long now_ms(void) {
    struct timeval tv;
    gettimeofday(&tv, NULL);
    return tv.tv_sec * 1000 + tv.tv_usec / 1000;
}

static void usage(const char* me) {
    fprintf(stderr, "\n");
    fprintf(stderr, " ---                                                       ---\n");
    fprintf(stderr, " \\  \\  mnemophonix - A simple audio fingerprinting system  \\  \\\n");
    fprintf(stderr, " O  O                                                      O  O\n");
    fprintf(stderr, "\n");
    fprintf(stderr, "Usage:\n");
    fprintf(stderr, "\n");
    fprintf(stderr, "%s index <input>\n", me);
    fprintf(stderr, "  Prints to stdout the index data generated for the given input file. Since\n");
    fprintf(stderr, "  the index file format is a text one, you can create a database containing\n");
    fprintf(stderr, "  multiple indexes like this:\n");
    fprintf(stderr, "\n");
    fprintf(stderr, "  $ %s index song1.mp3 > db\n", me);
    fprintf(stderr, "  $ %s index song2.wav >> db\n", me);
    fprintf(stderr, "  $ %s index movie.mp4 >> db\n", me);
    fprintf(stderr, "\n");
    fprintf(stderr, "%s search <input> <index>\n", me);
    fprintf(stderr, "  Looks for the given input file in the given index file\n");
    fprintf(stderr, "\n");
    fprintf(stderr, "The input file format that this code can process is 44100Hz 16-bit PCM.\n");
    fprintf(stderr, "If it is not the case, an attempt will be made to generate such a file using\n");
    fprintf(stderr, "ffmpeg. Because ** ffmpeg rocks **, you can use this program with pretty\n");
    fprintf(stderr, "much any audio or video file !\n");
    fprintf(stderr, "\n");
}

/* ffmpeg.c:9-32 upstream: artist= / title= / album= lines of an ffmetadata file (first occurrence wins). */
static void parse_metadata(const char* path, char** artist, char** track_title, char** album_title) {
    FILE* f = fopen(path, "r");
    if (!f) return;
    char line[1024];
    while (fgets(line, sizeof line, f)) {
        size_t len = strlen(line);
        if (len > 0 && line[len - 1] == '\n') line[len - 1] = '\0';
        if (artist && !*artist && !strncmp(line, "artist=", 7)) *artist = strdup(line + 7);
        if (track_title && !*track_title && !strncmp(line, "title=", 6)) *track_title = strdup(line + 6);
        if (album_title && !*album_title && !strncmp(line, "album=", 6)) *album_title = strdup(line + 6);
    }
    fclose(f);
}

/* Host input stage (ffmpeg.c:35-65 upstream): decode anything to a temporary 44.1 kHz PCM WAVE file with an external
 * ffmpeg, which also writes the stream's tags to an ffmetadata file (the reference's second output); artist, title
 * and album come from there.  Returns a malloc'd path or NULL. */
static char* decode_with_ffmpeg(const char* input, char** artist, char** track_title, char** album_title) {
    char tmpl[] = "/tmp/mnemophonixXXXXXX.wav";
    char meta[] = "/tmp/mnemophonixXXXXXX.meta";
    if (artist) *artist = NULL;
    if (track_title) *track_title = NULL;
    if (album_title) *album_title = NULL;
    int fd = mkstemps(tmpl, 4);
    if (fd < 0) return NULL;
    close(fd);
    fd = mkstemps(meta, 5);
    if (fd < 0) {
        remove(tmpl);
        return NULL;
    }
    close(fd);
    pid_t pid = fork();
    if (pid < 0) {
        remove(tmpl);
        remove(meta);
        return NULL;
    }
    if (pid == 0) {
        if (!freopen("/dev/null", "w", stderr)) _exit(127);
        execlp("ffmpeg", "ffmpeg", "-y", "-i", input, "-acodec", "pcm_s16le", "-ar", "44100", "-f", "wav", tmpl, "-f",
               "ffmetadata", meta, (char*)NULL);
        _exit(127);
    }
    int status = 0;
    if (waitpid(pid, &status, 0) < 0 || !WIFEXITED(status) || WEXITSTATUS(status) != 0) {
        remove(tmpl);
        remove(meta);
        return NULL;
    }
    parse_metadata(meta, artist, track_title, album_title);
    remove(meta);
    return strdup(tmpl);
}

/* generate_fingerprint with the reference's error texts and its ffmpeg fallback (main.c:57-84 upstream).
 * Returns 0 on success, 1 after printing the message. */
static int fingerprint_input(const char* input, struct signatures** fingerprint, char** artist, char** track_title,
                             char** album_title) {
    int res = generate_fingerprint(input, fingerprint, artist, track_title, album_title);
    switch (res) {
        case SUCCESS: return 0;
        case CANNOT_READ_FILE: fprintf(stderr, "Cannot read file '%s'\n", input); return 1;
        case MEMORY_ERROR: fprintf(stderr, "Memory allocation error\n"); return 1;
        case DECODING_ERROR: fprintf(stderr, "Cannot decode file '%s'\n", input); return 1;
        case FILE_TOO_SMALL: fprintf(stderr, "'%s' is too small to generate a fingerprint\n", input); return 1;
        case UNSUPPORTED_WAVE_FORMAT:
        case NOT_A_WAVE_FILE: {
            char* wav = decode_with_ffmpeg(input, artist, track_title, album_title);
            if (!wav) {
                fprintf(stderr, "'%s' is not a wave file and we could not convert it to one with fffmpeg\n", input);
                return 1;
            }
            res = generate_fingerprint(wav, fingerprint, NULL, NULL, NULL);
            remove(wav);
            free(wav);
            if (res == MEMORY_ERROR) { fprintf(stderr, "Memory allocation error\n"); return 1; }
            if (res == FILE_TOO_SMALL) { fprintf(stderr, "'%s' is too small to generate a fingerprint\n", input); return 1; }
            if (res != SUCCESS) { fprintf(stderr, "Cannot decode file '%s'\n", input); return 1; }
            return 0;
        }
        default: fprintf(stderr, "Cannot decode file '%s'\n", input); return 1;
    }
}

const char* mnemo_host_last_error(void);

static int report_to(FILE* out, const struct index* db, int best) {
    if (best < 0 && best != NO_MATCH_FOUND) {
        /* not a miss: the search itself failed (the reference can only run out of memory here) */
        fprintf(stderr, "Search failed (error %d): %s\n", best, mnemo_host_last_error());
        if (out != stdout) fprintf(out, "\nSearch failed (error %d)\n\n", best);
        return 2;
    }
    if (best < 0) {
        fprintf(out, "\nNo match found\n\n");
        return 1;
    }
    fprintf(out, "\nFound match: '%s'\n", db->entries[best]->filename);
    if (db->entries[best]->artist[0]) fprintf(out, "Artist: %s\n", db->entries[best]->artist);
    if (db->entries[best]->track_title[0]) fprintf(out, "Track title: %s\n", db->entries[best]->track_title);
    if (db->entries[best]->album_title[0]) fprintf(out, "Album title: %s\n", db->entries[best]->album_title);
    fprintf(out, "\n");
    return 0;
}

static int report(const struct index* db, int best) { return report_to(stdout, db, best); }

/* Extension: `mnemophonix serve <index> [socket]` — what the reference's live companion does (ears/main.m:132-139:
 * read_index and create_hash_tables once, then search again and again), as a resident process: the database and its
 * hash tables stay on the GPU and every request pays for one fingerprint and one search only.  Requests are input
 * paths, one per line, on stdin (answers on stdout) or, with a socket path, from clients of a UNIX stream socket
 * (answers on the connection); each answer is the block `search` prints, headed by `Searching '<path>'...`. */
/* Returns 1 if the stream asked the server to end (`shutdown`), 0 at its end or on `quit`. */
static int serve_stream(FILE* in, FILE* out, struct index* db, struct lsh* lsh) {
    char line[4096];
    while (fgets(line, sizeof line, in)) {
        size_t len = strlen(line);
        while (len && (line[len - 1] == '\n' || line[len - 1] == '\r')) line[--len] = '\0';
        if (!len) continue;
        if (!strcmp(line, "quit")) break;
        if (!strcmp(line, "shutdown")) return 1;
        struct signatures* fp = NULL;
        fprintf(out, "Searching '%s'...\n", line);
        if (fingerprint_input(line, &fp, NULL, NULL, NULL)) {
            fprintf(out, "\nCannot fingerprint '%s'\n\n", line);
            fflush(out);
            continue;
        }
        long t0 = now_ms();
        int best = search(fp, db, lsh, 0);
        fprintf(out, "(Search took %ld ms)\n", now_ms() - t0);
        report_to(out, db, best);
        fflush(out);
        free_signatures(fp);
    }
    return 0;
}

static int serve(const char* index, const char* sock_path) {
    struct index* db = NULL;
    printf("Loading database %s...\n", index);
    long t0 = now_ms();
    int res = read_index(index, &db);
    if (res != SUCCESS) {
        if (res == CANNOT_READ_FILE) fprintf(stderr, "Cannot read file '%s'\n", index);
        else if (res == MEMORY_ERROR) fprintf(stderr, "Memory allocation error\n");
        else fprintf(stderr, "Cannot decode file '%s'\n", index);
        return 1;
    }
    long t1 = now_ms();
    printf("(raw database loading took %ld ms)\n", t1 - t0);
    struct lsh* lsh = create_hash_tables(db);
    printf("(lsh index building took %ld ms)\n", now_ms() - t1);
    if (!lsh) {
        fprintf(stderr, "Memory allocation error (%s)\n", mnemo_host_last_error());
        return 1;
    }
    if (!sock_path) {
        printf("Ready\n");
        fflush(stdout);
        serve_stream(stdin, stdout, db, lsh);
    } else {
        struct sockaddr_un addr;
        memset(&addr, 0, sizeof addr);
        addr.sun_family = AF_UNIX;
        if (strlen(sock_path) >= sizeof addr.sun_path) {
            fprintf(stderr, "Socket path too long\n");
            return 1;
        }
        strcpy(addr.sun_path, sock_path);
        unlink(sock_path);
        int srv = socket(AF_UNIX, SOCK_STREAM, 0);
        if (srv < 0 || bind(srv, (struct sockaddr*)&addr, sizeof addr) != 0 || listen(srv, 16) != 0) {
            fprintf(stderr, "Cannot listen on '%s'\n", sock_path);
            return 1;
        }
        signal(SIGPIPE, SIG_IGN);
        printf("Ready on %s\n", sock_path);
        fflush(stdout);
        for (;;) {   /* one client at a time: the GPU is the shared resource anyway; the line `shutdown` ends the server */
            int c = accept(srv, NULL, NULL);
            if (c < 0) break;
            FILE* in = fdopen(c, "r");
            FILE* out = fdopen(dup(c), "w");
            const int stop = (in && out) ? serve_stream(in, out, db, lsh) : 0;
            if (in) fclose(in);
            if (out) fclose(out);
            if (stop) break;
        }
        close(srv);
        unlink(sock_path);
    }
    free_hash_tables(lsh);
    free_index(db);
    return 0;
}

int main(int argc, char* argv[]) {
    /* `search` takes <input> <index> as upstream; with more arguments (an extension) every argument but the last
     * is an input and the last one the index: the database is loaded and its tables are built once. */
    if (argc >= 3 && !strcmp(argv[1], "serve")) return serve(argv[2], argc > 3 ? argv[3] : NULL);
    if (argc < 2 || (strcmp(argv[1], "index") && strcmp(argv[1], "search")) || argc < 3 ||
        (!strcmp(argv[1], "search") && argc < 4)) {
        usage(argv[0]);
        return 1;
    }
    char* input = argv[2];
    struct signatures* fingerprint = NULL;
    char *artist = NULL, *track_title = NULL, *album_title = NULL;
    if (fingerprint_input(input, &fingerprint, &artist, &track_title, &album_title)) return 1;

    if (!strcmp(argv[1], "index")) {
        save(stdout, fingerprint, input, artist, track_title, album_title);
        /* Extension (the reference ignores further arguments): `index a b c ...` appends one entry per
         * input, in order, paying for the CUDA context once.  A failing input ends the run with status 1
         * after the entries already written, so the output stays a valid database prefix. */
        for (int i = 3; i < argc; i++) {
            free_signatures(fingerprint);
            free(artist);
            free(track_title);
            free(album_title);
            fingerprint = NULL;
            artist = track_title = album_title = NULL;
            if (fingerprint_input(argv[i], &fingerprint, &artist, &track_title, &album_title)) return 1;
            fflush(stdout);
            save(stdout, fingerprint, argv[i], artist, track_title, album_title);
        }
        return 0;
    }

    const char* index = argv[argc - 1];
    struct index* db = NULL;
    printf("Loading database %s...\n", index);
    long t0 = now_ms();
    int res = read_index(index, &db);
    if (res != SUCCESS) {
        if (res == CANNOT_READ_FILE) fprintf(stderr, "Cannot read file '%s'\n", index);
        else if (res == MEMORY_ERROR) fprintf(stderr, "Memory allocation error\n");
        else fprintf(stderr, "Cannot decode file '%s'\n", index);
        return 1;
    }
    long t1 = now_ms();
    printf("(raw database loading took %ld ms)\n", t1 - t0);
    struct lsh* lsh = create_hash_tables(db);
    long t2 = now_ms();
    printf("(lsh index building took %ld ms)\n", t2 - t1);
    if (!lsh) {
        fprintf(stderr, "Memory allocation error (%s)\n", mnemo_host_last_error());
        return 1;
    }
    printf("Searching...\n");
    int best = search(fingerprint, db, lsh, 0);
    long t3 = now_ms();
    printf("(Search took %ld ms)\n", t3 - t2);
    int status = report(db, best);
    /* further inputs (extension): same block per input, headed by its name; exit status 1 if any input failed or
     * found nothing */
    for (int i = 3; i < argc - 1; i++) {
        free_signatures(fingerprint);
        fingerprint = NULL;
        printf("Searching '%s'...\n", argv[i]);
        if (fingerprint_input(argv[i], &fingerprint, NULL, NULL, NULL)) {
            status = 1;
            continue;
        }
        long t4 = now_ms();
        best = search(fingerprint, db, lsh, 0);
        printf("(Search took %ld ms)\n", now_ms() - t4);
        status |= report(db, best);
    }
    return status;
}
