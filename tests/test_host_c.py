"""The drop-in C library (libmnemophonix.so: reference-named entry points over the CUDA C ABI) and CLI.
CPU part: text database I/O and WAV container handling against the reference's own functions, error codes.
GPU part: generate_fingerprint / create_hash_tables / get_matches / search / the command line."""
import ctypes as C
import os
import struct
import subprocess
import tempfile

import numpy as np
import pytest

from mnemophonix_b200 import synth
from oracle import ref as refmod

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "mnemophonix_b200", "libmnemophonix.so")
CLI = os.path.join(ROOT, "mnemophonix_b200", "mnemophonix")


@pytest.fixture(scope="module")
def host():
    lib = C.CDLL(LIB)
    lib.read_index.argtypes = [C.c_char_p, C.POINTER(C.POINTER(refmod.Index))]
    lib.free_index.argtypes = [C.POINTER(refmod.Index)]
    lib.save.argtypes = [C.c_void_p, C.POINTER(refmod.Signatures), C.c_char_p, C.c_char_p, C.c_char_p, C.c_char_p]
    lib.generate_fingerprint.argtypes = [C.c_char_p, C.POINTER(C.POINTER(refmod.Signatures)), C.POINTER(C.c_char_p),
                                         C.POINTER(C.c_char_p), C.POINTER(C.c_char_p)]
    lib.generate_fingerprint_from_samples.argtypes = [C.c_void_p, C.c_uint, C.POINTER(C.POINTER(refmod.Signatures))]
    lib.free_signatures.argtypes = [C.POINTER(refmod.Signatures)]
    lib.create_hash_tables.argtypes = [C.POINTER(refmod.Index)]
    lib.create_hash_tables.restype = C.c_void_p
    lib.free_hash_tables.argtypes = [C.c_void_p]
    lib.get_matches.argtypes = [C.c_void_p, C.c_void_p, C.POINTER(C.POINTER(refmod.SignatureList))]
    lib.free_signature_list.argtypes = [C.POINTER(refmod.SignatureList)]
    lib.search.argtypes = [C.POINTER(refmod.Signatures), C.POINTER(refmod.Index), C.c_void_p, C.c_int]
    return lib


def db_text(tracks, names=None):
    out = []
    for i, s in enumerate(tracks):
        name = names[i] if names else f"track{i}.wav"
        out.append(f"{name}\nartist{i}\n\nalbum {i}\n{len(s)}\n")
        out.append("".join("".join(f"{b:02x}" for b in row) + "\n" for row in s))
    return "".join(out)


def save_with(lib, sigs, name, artist, title, album):
    libc = C.CDLL(None)
    libc.fopen.restype = C.c_void_p
    libc.fopen.argtypes = [C.c_char_p, C.c_char_p]
    libc.fclose.argtypes = [C.c_void_p]
    p = tempfile.mktemp()
    f = libc.fopen(p.encode(), b"w")
    sigs = np.ascontiguousarray(sigs, np.uint8)
    s = refmod.Signatures(len(sigs), C.cast(sigs.ctypes.data, C.POINTER(C.c_uint8 * 100)))
    lib.save.argtypes = [C.c_void_p, C.POINTER(refmod.Signatures), C.c_char_p, C.c_char_p, C.c_char_p, C.c_char_p]
    lib.save(f, C.byref(s), name, artist, title, album)
    libc.fclose(f)
    data = open(p, "rb").read()
    os.remove(p)
    return data


def entries_of(pidx):
    out = []
    for i in range(pidx.contents.n_entries):
        e = pidx.contents.entries[i].contents
        n = e.signatures.contents.n_signatures
        sig = np.ctypeslib.as_array(C.cast(e.signatures.contents.signatures, C.POINTER(C.c_uint8)), (n, 100)).copy() \
            if n else np.zeros((0, 100), np.uint8)
        out.append((e.filename, e.artist, e.track_title, e.album_title, sig))
    return out


def test_save_is_byte_identical_to_reference(host, ref):
    rng = np.random.RandomState(0)
    sigs = rng.randint(0, 256, (37, 100)).astype(np.uint8)
    for meta in ((b"a b.wav", b"art", b"", None), (b"x", None, None, None)):
        assert save_with(host, sigs, *meta) == save_with(ref.lib, sigs, *meta)
    assert save_with(host, sigs[:0], b"e", None, None, None) == save_with(ref.lib, sigs[:0], b"e", None, None, None)


def test_read_index_matches_reference(host, ref):
    rng = np.random.RandomState(1)
    tracks = [rng.randint(0, 256, (n, 100)).astype(np.uint8) for n in (5, 0, 17, 1)]
    text = db_text(tracks).replace("0a", "0A", 3)          # upper-case hex is accepted (fingerprintio.c:144)
    p = tempfile.mktemp()
    open(p, "w").write(text)
    try:
        mine = C.POINTER(refmod.Index)()
        assert host.read_index(p.encode(), C.byref(mine)) == 0
        rc, theirs = ref.load_index(p)
        assert rc == 0
        a, b = entries_of(mine), entries_of(theirs)
        assert len(a) == len(b) == 4
        for x, y in zip(a, b):
            assert x[:4] == y[:4] and np.array_equal(x[4], y[4])
        host.free_index(mine)
        ref.lib.free_index(theirs)
    finally:
        os.remove(p)


def test_large_index_takes_the_threaded_parser_and_matches_reference(host, ref):
    """A database big enough for the multi-threaded fast path of read_index (mapped file, entries located by one walk
    over the headers, signature blocks decoded in parallel): same entries as the reference's parser; a corrupt byte
    deep inside, a short line and a truncated file are rejected with the reference's code by the sequential parser."""
    import binascii
    rng = np.random.RandomState(7)
    tracks = [rng.randint(0, 256, (int(n), 100)).astype(np.uint8) for n in rng.randint(0, 700, 700)]
    assert sum(len(t) for t in tracks) > 200000
    parts = []
    for i, t in enumerate(tracks):
        hexed = binascii.hexlify(t.tobytes()).decode()
        parts.append(f"dir/track {i}.wav\nartist {i % 7}\n\nalbum {i % 3}\n{len(t)}\n")
        parts.append("".join(hexed[200 * k:200 * k + 200] + "\n" for k in range(len(t))))
    text = "".join(parts)
    p = tempfile.mktemp()
    try:
        open(p, "w").write(text)
        mine = C.POINTER(refmod.Index)()
        assert host.read_index(p.encode(), C.byref(mine)) == 0
        assert mine.contents.n_entries == len(tracks)
        got = entries_of(mine)
        for i in (0, 1, 2, 99, 350, 698, 699):                      # spot-check against the reference structures below
            assert got[i][0] == f"dir/track {i}.wav".encode() and np.array_equal(got[i][4], tracks[i])
        assert all(np.array_equal(g[4], t) for g, t in zip(got, tracks))
        rc, theirs = ref.load_index(p)
        assert rc == 0
        want = entries_of(theirs)
        assert all(x[:4] == y[:4] and np.array_equal(x[4], y[4]) for x, y in zip(got, want)) and len(got) == len(want)
        host.free_index(mine)
        ref.lib.free_index(theirs)
        mid = len(text) // 2
        line_start = text.rfind("\n", 0, mid) + 1
        while len(text[line_start:text.find("\n", line_start)]) != 200:   # find a signature line near the middle
            line_start = text.find("\n", line_start) + 1
        for bad in (text[:line_start + 10] + "g" + text[line_start + 11:],          # not a hex digit
                    text[:line_start + 10] + text[line_start + 11:],                # one character short
                    text[:mid]):                                                   # cut in the middle
            open(p, "w").write(bad)
            mine = C.POINTER(refmod.Index)()
            assert host.read_index(p.encode(), C.byref(mine)) == -3
            assert ref.load_index(p)[0] == -3
    finally:
        if os.path.exists(p):
            os.remove(p)


def test_over_long_first_line_is_kept_like_the_reference(host, ref):
    """fingerprintio.c:78-80 only checks the first line of an entry for the end of the file: a file name longer than
    the 1024-byte line buffer is not an error, the entry simply starts with the first 1023 characters and the rest of
    the line is read as the artist.  Same entries, or the same error, as the reference."""
    rng = np.random.RandomState(4)
    sig = rng.randint(0, 256, (2, 100)).astype(np.uint8)
    hexed = "".join("".join(f"{b:02x}" for b in row) + "\n" for row in sig)
    for name_len in (1500, 1023):
        # what follows the first 1023 characters is the artist line; title, album and count complete a valid entry
        # (anything else sends the reference into its uninitialised free, fingerprintio.c:117-119)
        text = "n" * name_len + "\ntitle\nalbum\n2\n" + hexed
        p = tempfile.mktemp()
        open(p, "w").write(text)
        try:
            mine = C.POINTER(refmod.Index)()
            rc_mine = host.read_index(p.encode(), C.byref(mine))
            rc_ref, theirs = ref.load_index(p)
            assert rc_mine == rc_ref == 0, name_len
            a, b = entries_of(mine), entries_of(theirs)
            assert len(a) == len(b) == 1 and all(x[:4] == y[:4] and np.array_equal(x[4], y[4]) for x, y in zip(a, b))
            assert len(a[0][0]) == 1023 and a[0][1] == b"n" * (name_len - 1023)
            host.free_index(mine)
            ref.lib.free_index(theirs)
        finally:
            os.remove(p)


@pytest.mark.parametrize("mutation,code", [
    (lambda t: t[:-50], -3),                               # truncated last signature line
    (lambda t: t.replace("\n", "", 1), "ours:-3"),         # count line is not a number: DECODING_ERROR by the source, but the
                                                           # reference then frees an uninitialised pointer (fingerprintio.c:117-119)
    (lambda t: t[:t.rfind("\n", 0, len(t) - 1) + 1] + "zz" + t[t.rfind("\n", 0, len(t) - 1) + 3:], -3),   # bad hex
    (lambda t: t + "dangling\n", -3),                      # entry cut after its first line
])
def test_read_index_error_codes_match_reference(host, ref, mutation, code):
    rng = np.random.RandomState(2)
    text = mutation(db_text([rng.randint(0, 256, (3, 100)).astype(np.uint8)]))
    p = tempfile.mktemp()
    open(p, "w").write(text)
    try:
        mine = C.POINTER(refmod.Index)()
        if isinstance(code, str):
            assert host.read_index(p.encode(), C.byref(mine)) == int(code.split(":")[1])
        else:
            rc_ref, _ = ref.load_index(p)
            assert host.read_index(p.encode(), C.byref(mine)) == rc_ref == code
    finally:
        os.remove(p)
    assert host.read_index(b"/nonexistent/db", C.byref(mine)) == -2


def _wav(fmt_tag=1, rate=44100, bits=16, ch=1, fmt_size=16, head=b"RIFF", form=b"WAVE", n=20000, extra=b"", cut=0):
    pcm = synth.synth_track(3, n / 44100.0, ch)[:n]
    data = pcm.tobytes()
    fmt = b"fmt " + struct.pack("<IhHIIHH", fmt_size, fmt_tag, ch, rate, rate * ch * bits // 8, ch * bits // 8, bits)
    fmt += b"\0" * (fmt_size - 16)
    body = form + fmt + extra + b"data" + struct.pack("<I", len(data)) + data
    blob = head + struct.pack("<I", len(body)) + body
    return blob[:len(blob) - cut] if cut else blob


@pytest.mark.parametrize("kwargs,code", [
    (dict(head=b"RIFX"), -6), (dict(form=b"AVI "), -6),                       # NOT_A_WAVE_FILE
    (dict(fmt_tag=3), -4), (dict(rate=48000), -4), (dict(bits=8), -4), (dict(fmt_size=18), -4),   # UNSUPPORTED
    (dict(n=4000), -5),                                                       # FILE_TOO_SMALL (500 samples at 5512 Hz)
    (dict(cut=1000), "ours:-3"),   # data chunk shorter than declared: read_samples returns DECODING_ERROR, after which the
                                   # reference frees an uninitialised pointer (fingerprinting.c:43) and crashes
    (dict(n=100003, extra=b"junk" + struct.pack("<I", 6) + b"abcdef"), None),   # skipped chunk, needs a GPU
])
def test_wav_error_codes_match_reference(host, ref, kwargs, code):
    p = tempfile.mktemp(suffix=".wav")
    open(p, "wb").write(_wav(**kwargs))
    try:
        if isinstance(code, str):
            psig = C.POINTER(refmod.Signatures)()
            assert host.generate_fingerprint(p.encode(), C.byref(psig), None, None, None) == int(code.split(":")[1])
            return
        rc_ref = ref.fingerprint_wav(p)[0]
        if code is None:
            assert rc_ref == 0
            return
        psig = C.POINTER(refmod.Signatures)()
        rc = host.generate_fingerprint(p.encode(), C.byref(psig), None, None, None)
        assert rc == rc_ref == code
    finally:
        os.remove(p)
    psig = C.POINTER(refmod.Signatures)()
    assert host.generate_fingerprint(b"/nonexistent.wav", C.byref(psig), None, None, None) == -2


def test_cli_usage_and_exit_status():
    r = subprocess.run([CLI], capture_output=True, text=True)
    assert r.returncode == 1 and "mnemophonix - A simple audio fingerprinting system" in r.stderr
    r = subprocess.run([CLI, "search", "x.wav"], capture_output=True, text=True)
    assert r.returncode == 1 and "Usage:" in r.stderr
    r = subprocess.run([CLI, "index", "/nonexistent.wav"], capture_output=True, text=True)
    assert r.returncode == 1 and r.stderr.strip().endswith("Cannot read file '/nonexistent.wav'")


# ------------------------------------------------------------------------------ GPU ----

@pytest.mark.gpu
def test_generate_fingerprint_and_metadata(host, ref):
    pcm = synth.synth_track(77, 9.0, 2)
    p = tempfile.mktemp(suffix=".wav")
    synth.write_wav(p, pcm, {"IART": "an artist", "INAM": "a title", "IPRD": "an album"})
    try:
        psig = C.POINTER(refmod.Signatures)()
        a, t, al = C.c_char_p(), C.c_char_p(), C.c_char_p()
        assert host.generate_fingerprint(p.encode(), C.byref(psig), C.byref(a), C.byref(t), C.byref(al)) == 0
        rc, want, ra, rt, ral = ref.fingerprint_wav(p)
        got = np.ctypeslib.as_array(C.cast(psig.contents.signatures, C.POINTER(C.c_uint8)),
                                    (psig.contents.n_signatures, 100)).copy()
        assert np.array_equal(got, want) and (a.value, t.value, al.value) == (ra, rt, ral)
        host.free_signatures(psig)
    finally:
        os.remove(p)


@pytest.mark.gpu
def test_streamed_ingest_of_a_multi_piece_file(host, ref):
    # 250 s of stereo = 44 MB: two 32 MB pieces through the pinned staging pair (host/fingerprinting.c)
    pcm = synth.synth_track(78, 250.0, 2)
    p = tempfile.mktemp(suffix=".wav")
    synth.write_wav(p, pcm)
    try:
        psig = C.POINTER(refmod.Signatures)()
        assert host.generate_fingerprint(p.encode(), C.byref(psig), None, None, None) == 0
        rc, want, *_ = ref.fingerprint_wav(p)
        got = np.ctypeslib.as_array(C.cast(psig.contents.signatures, C.POINTER(C.c_uint8)),
                                    (psig.contents.n_signatures, 100)).copy()
        assert rc == 0 and np.array_equal(got, want)
        host.free_signatures(psig)
        # the same file cut short inside its second piece: read_samples' DECODING_ERROR (wav.c:381-386)
        with open(p, "r+b") as f:
            f.truncate(os.path.getsize(p) - 5_000_001)
        psig = C.POINTER(refmod.Signatures)()
        assert host.generate_fingerprint(p.encode(), C.byref(psig), None, None, None) == -3
    finally:
        os.remove(p)


@pytest.mark.gpu
def test_cli_index_and_search_round_trip(host, ref):
    d = tempfile.mkdtemp()
    names = []
    db = b""
    for i in range(3):
        pcm = synth.synth_track(300 + i, 25.0, 1 + i % 2)
        path = os.path.join(d, f"song{i}.wav")
        synth.write_wav(path, pcm, {"IART": f"artist {i}", "INAM": f"title {i}"})
        names.append(path)
        r = subprocess.run([CLI, "index", path], capture_output=True)
        assert r.returncode == 0 and b"Generated" in r.stderr
        rc, sigs, a, t, al = ref.fingerprint_wav(path)
        assert r.stdout == save_with(ref.lib, sigs, path.encode(), a, t, al)      # byte-identical entry text
        db += r.stdout
    r = subprocess.run([CLI, "index"] + names, capture_output=True)     # multi-file extension: same bytes, one process
    assert r.returncode == 0 and r.stdout == db
    dbpath = os.path.join(d, "db")
    open(dbpath, "wb").write(db)
    clip = os.path.join(d, "clip.wav")
    synth.write_wav(clip, synth.synth_track(301, 25.0, 2)[7 * 44100 + 999: 12 * 44100 + 999])
    r = subprocess.run([CLI, "search", clip, dbpath], capture_output=True, text=True)
    assert r.returncode == 0
    assert f"Found match: '{names[1]}'" in r.stdout and "Artist: artist 1" in r.stdout and "Track title: title 1" in r.stdout
    for line in ("Loading database", "(raw database loading took", "(lsh index building took", "Searching...", "(Search took"):
        assert line in r.stdout
    noise = os.path.join(d, "noise.wav")
    synth.write_wav(noise, (np.random.RandomState(0).standard_normal((44100 * 5, 1)) * 3000).astype(np.int16))
    r = subprocess.run([CLI, "search", noise, dbpath], capture_output=True, text=True)
    assert r.returncode == 1 and "No match found" in r.stdout
    # several inputs per invocation (extension): one database load, one block per input, status 1 if any found nothing
    clip2 = os.path.join(d, "clip2.wav")
    synth.write_wav(clip2, synth.synth_track(302, 25.0, 1)[9 * 44100 + 5: 14 * 44100 + 5])
    r = subprocess.run([CLI, "search", clip, noise, clip2, dbpath], capture_output=True, text=True)
    assert r.returncode == 1 and r.stdout.count("Loading database") == 1
    blocks = r.stdout.split("Searching")
    assert f"Found match: '{names[1]}'" in blocks[1] and "No match found" in blocks[2] and f"Found match: '{names[2]}'" in blocks[3]
    r = subprocess.run([CLI, "search", clip, clip2, dbpath], capture_output=True, text=True)
    assert r.returncode == 0 and r.stdout.count("Found match") == 2

    # the C API on the same database: create_hash_tables / get_matches / search against the reference's
    mine = C.POINTER(refmod.Index)()
    assert host.read_index(dbpath.encode(), C.byref(mine)) == 0
    rc, theirs = ref.load_index(dbpath)
    lsh_mine = host.create_hash_tables(mine)
    lsh_ref = ref.lib.create_hash_tables(theirs)
    assert C.cast(lsh_mine, C.POINTER(C.c_uint)).contents.value == C.cast(lsh_ref, C.POINTER(C.c_uint)).contents.value
    rc, q, *_ = ref.fingerprint_wav(clip)
    s = refmod.Signatures(len(q), C.cast(q.ctypes.data, C.POINTER(C.c_uint8 * 100)))
    assert host.search(C.byref(s), mine, lsh_mine, 0) == ref.search(q, theirs, lsh_ref) == 1
    plist = C.POINTER(refmod.SignatureList)()
    n = host.get_matches(lsh_mine, q[2].ctypes.data, C.byref(plist))
    got, node = [], plist
    while node:
        got.append((node.contents.entry_index, node.contents.signature_index))
        node = node.contents.next
    assert n == len(got) and got == ref.get_matches(lsh_ref, q[2])
    host.free_signature_list(plist)
    host.free_hash_tables(lsh_mine)
    host.free_index(mine)


@pytest.mark.gpu
def test_resident_search_server(oracle):
    """`mnemophonix serve <index> [socket]` (extension; the ears loop of the reference as a process: database and hash
    tables built once and kept on the GPU, one fingerprint + search per request): requests on stdin, and from two
    consecutive clients of a UNIX socket; every answer block equals what `mnemophonix search` prints for that input."""
    import socket
    import time
    d = tempfile.mkdtemp()
    names, entries = [], []
    for i in range(4):
        pcm = synth.synth_track(820 + i, 16.0, 1 + i % 2)
        path = os.path.join(d, f"s{i}.wav")
        synth.write_wav(path, pcm, {"IART": f"artist {i}"})
        names.append(path)
        entries.append(oracle.fingerprint_pcm(pcm)[1])
    dbpath = os.path.join(d, "db")
    open(dbpath, "w").write(db_text(entries, names))
    clips = []
    for i in (2, 0):
        c = os.path.join(d, f"clip{i}.wav")
        synth.write_wav(c, synth.synth_track(820 + i, 16.0, 1 + i % 2)[4 * 44100 + 55: 9 * 44100 + 55])
        clips.append(c)
    noise = os.path.join(d, "noise.wav")
    synth.write_wav(noise, (np.random.RandomState(1).standard_normal((44100 * 5, 1)) * 3000).astype(np.int16))

    def block_of(path):       # what the one-shot CLI prints after its timing lines
        r = subprocess.run([CLI, "search", path, dbpath], capture_output=True, text=True)
        return r.stdout[r.stdout.index("(Search took"):].split("\n", 1)[1]

    want = {p: block_of(p) for p in clips + [noise]}
    assert f"Found match: '{names[2]}'" in want[clips[0]] and "No match found" in want[noise]
    r = subprocess.run([CLI, "serve", dbpath], input="\n".join([clips[0], noise, "/nonexistent.wav", clips[1]]) + "\n",
                       capture_output=True, text=True, timeout=300)
    assert r.returncode == 0 and r.stdout.count("Loading database") == 1 and "Ready" in r.stdout
    parts = r.stdout.split("Searching '")[1:]
    assert [p.split("'...")[0] for p in parts] == [clips[0], noise, "/nonexistent.wav", clips[1]]
    for p, path in zip(parts, [clips[0], noise, None, clips[1]]):
        if path is None:
            assert "Cannot fingerprint" in p
        else:
            assert p.split("\n", 2)[2] == want[path], (p, want[path])
    sock = os.path.join(d, "mnemo.sock")
    srv = subprocess.Popen([CLI, "serve", dbpath, sock], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
    try:
        for _ in range(600):
            if os.path.exists(sock):
                break
            time.sleep(0.1)
        for path in (clips[1], noise):                      # two clients, one after the other
            c = socket.socket(socket.AF_UNIX)
            c.connect(sock)
            c.sendall((path + "\nquit\n").encode())
            data = b""
            while True:
                chunk = c.recv(65536)
                if not chunk:
                    break
                data += chunk
            c.close()
            text = data.decode()
            assert text.startswith(f"Searching '{path}'...") and text.split("\n", 2)[2] == want[path]
        c = socket.socket(socket.AF_UNIX)
        c.connect(sock)
        c.sendall(b"shutdown\n")
        c.close()
        assert srv.wait(timeout=60) == 0
    finally:
        if srv.poll() is None:
            srv.kill()


_VERBOSE_SCRIPT = """
import sys, ctypes as C
sys.path.insert(0, %r)
import numpy as np
from oracle import ref as refmod
which, dbpath, qpath = sys.argv[1:4]
q = np.load(qpath)
r = refmod.Ref()
if which == "ref":
    lib = r.lib
    rc, idx = r.load_index(dbpath)
else:
    lib = C.CDLL(%r)
    idx = C.POINTER(refmod.Index)()
    lib.read_index.argtypes = [C.c_char_p, C.POINTER(C.POINTER(refmod.Index))]
    assert lib.read_index(dbpath.encode(), C.byref(idx)) == 0
lib.create_hash_tables.argtypes = [C.POINTER(refmod.Index)]
lib.create_hash_tables.restype = C.c_void_p
lib.search.argtypes = [C.POINTER(refmod.Signatures), C.POINTER(refmod.Index), C.c_void_p, C.c_int]
lsh = lib.create_hash_tables(idx)
for k in range(len(q)):
    qq = np.ascontiguousarray(q[k])
    s = refmod.Signatures(len(qq), C.cast(qq.ctypes.data, C.POINTER(C.c_uint8 * 100)))
    best = lib.search(C.byref(s), idx, lsh, 1)
    C.CDLL(None).fflush(None)
    print("RETURNED", best, flush=True)
"""


@pytest.mark.gpu
def test_verbose_search_prints_the_reference_lines(oracle):
    """search(..., verbose) (search.c:175-188): the ten `average_score = ...` lines of the sorted scores[] array (matched
    entries in the merge sort's order, then unmatched entries in index order), the match line and the rule — the
    whole stdout of the reference, byte for byte, for a 14-entry database (more than ten entries, so the window
    truncates), a matching clip, a clip matching nothing and a query shorter than the thresholds."""
    d = tempfile.mkdtemp()
    pcms = [synth.synth_track(700 + i, 14.0 + (i % 3), 1) for i in range(14)]
    sigs = [oracle.fingerprint_pcm(p)[1] for p in pcms]
    dbpath = os.path.join(d, "db")
    open(dbpath, "w").write(db_text(sigs))
    qs = [oracle.fingerprint_pcm(pcms[11][3 * 44100 + 77: 8 * 44100 + 77])[1],
          np.random.RandomState(3).randint(0, 255, (34, 100)).astype(np.uint8),
          oracle.fingerprint_pcm(pcms[2][5 * 44100: 8 * 44100 + 4000])[1][:3]]
    width = max(len(q) for q in qs)
    qpath = os.path.join(d, "q.npy")
    np.save(qpath, np.stack([np.concatenate([q, np.repeat(q[-1:], width - len(q), 0)]) for q in qs]))
    outs = {}
    for which in ("ref", "ours"):
        r = subprocess.run([os.sys.executable, "-c", _VERBOSE_SCRIPT % (ROOT, LIB), which, dbpath, qpath], capture_output=True,
                           text=True, timeout=600)
        assert r.returncode == 0, r.stderr[-2000:]
        outs[which] = r.stdout
    assert "average_score = " in outs["ref"] and "*** match = " in outs["ref"] and "RETURNED 11" in outs["ref"]
    assert outs["ours"] == outs["ref"]


@pytest.mark.gpu
def test_concurrent_callers_share_one_context(host, oracle):
    """SURVEY.md §8(b) threading row: the reference's API is re-entrant after warm-up (search / LSH are read-only
    on shared index + lsh), so the drop-in must accept calls from any thread at once.  Eight threads index
    different WAV files and search one shared database simultaneously (ctypes releases the GIL during the calls);
    every result must equal the serial one."""
    import threading
    d = tempfile.mkdtemp()
    paths, want = [], []
    for i in range(8):
        pcm = synth.synth_track(900 + i, 12.0 + 2.5 * i, 1 + i % 2)
        p = os.path.join(d, f"t{i}.wav")
        synth.write_wav(p, pcm)
        paths.append(p)
        rc, s = oracle.fingerprint_pcm(pcm)
        assert rc == 0
        want.append(s)
    dbpath = os.path.join(d, "db")
    open(dbpath, "w").write(db_text(want, paths))
    idx = C.POINTER(refmod.Index)()
    assert host.read_index(dbpath.encode(), C.byref(idx)) == 0
    lsh = host.create_hash_tables(idx)
    assert lsh
    queries = []
    for i in range(8):
        clip = synth.synth_track(900 + i, 12.0 + 2.5 * i, 1 + i % 2)[3 * 44100 + 321: 8 * 44100 + 321]
        rc, q = oracle.fingerprint_pcm(clip)
        assert rc == 0 and len(q)
        queries.append(np.ascontiguousarray(q))
    odb = oracle.make_db(want)
    best = [oracle.search(odb, q)[0] for q in queries]
    n_nodes = [len(oracle.get_matches(odb, q[0])) for q in queries]
    assert best == list(range(8))          # every excerpt identifies its own track (serial, CPU oracle)
    errors = []

    def worker(i):
        try:
            for rep in range(3):
                psig = C.POINTER(refmod.Signatures)()
                rc = host.generate_fingerprint(paths[i].encode(), C.byref(psig), None, None, None)
                assert rc == 0, rc
                got = np.ctypeslib.as_array(C.cast(psig.contents.signatures, C.POINTER(C.c_uint8)),
                                            (psig.contents.n_signatures, 100)).copy()
                host.free_signatures(psig)
                assert np.array_equal(got, want[i]), f"thread {i}: signatures differ"
                q = queries[(i + rep) % 8]
                s = refmod.Signatures(len(q), C.cast(q.ctypes.data, C.POINTER(C.c_uint8 * 100)))
                assert host.search(C.byref(s), idx, lsh, 0) == best[(i + rep) % 8]
                plist = C.POINTER(refmod.SignatureList)()
                n = host.get_matches(lsh, q[0].ctypes.data, C.byref(plist))
                assert n == n_nodes[(i + rep) % 8]
                host.free_signature_list(plist)
        except BaseException as e:   # noqa: BLE001 - collected and re-raised on the main thread
            errors.append(e)

    threads = [threading.Thread(target=worker, args=(i,)) for i in range(8)]
    for t in threads:
        t.start()
    for t in threads:
        t.join()
    host.free_hash_tables(lsh)
    host.free_index(idx)
    if errors:
        raise errors[0]
