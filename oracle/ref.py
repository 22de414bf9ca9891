"""ctypes harness around the UNMODIFIED reference library built into oracle/_ref/.

TEST INFRASTRUCTURE ONLY (tests/, __graft_entry__.smoke(), bench.py's cpu_baseline /
--impl reference).  Never imported by the product package.

Protocol: the reference fills its FFT bit-reversal table, band-edge
table and Hann window lazily and racily on first use from 8 pthreads
(fft.c:71-75, logbins.c:59-62, hannwindow.c:15-21, spectralimages.c:160-178), so the
harness pre-warms those three tables single-threaded right after dlopen.  With that the
library is deterministic and -O0 == -O2 byte for byte.
"""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))

SIG_LEN = 100
IMG_ELEMS = 128 * 32
RAWFP_BYTES = 1024


class Signatures(C.Structure):
    _fields_ = [("n_signatures", C.c_uint), ("signatures", C.POINTER(C.c_uint8 * SIG_LEN))]


class SpectralImages(C.Structure):
    _fields_ = [("n_images", C.c_uint), ("images", C.POINTER(C.c_float * IMG_ELEMS))]


class RawFingerprint(C.Structure):
    _pack_ = 1
    _fields_ = [("is_silence", C.c_char), ("bit_array", C.c_uint8 * RAWFP_BYTES)]


class RawFingerprints(C.Structure):
    _fields_ = [("size", C.c_uint), ("fingerprints", C.POINTER(RawFingerprint))]


class IndexEntry(C.Structure):
    _fields_ = [("filename", C.c_char_p), ("artist", C.c_char_p), ("track_title", C.c_char_p),
                ("album_title", C.c_char_p), ("signatures", C.POINTER(Signatures))]


class Index(C.Structure):
    _fields_ = [("n_entries", C.c_uint), ("entries", C.POINTER(C.POINTER(IndexEntry)))]


class SignatureList(C.Structure):
    pass


SignatureList._fields_ = [("entry_index", C.c_uint), ("signature_index", C.c_uint),
                          ("next", C.POINTER(SignatureList))]


def ref_path(variant: str = "O2") -> str:
    return os.path.join(HERE, "_ref", f"libmnemophonix_{variant}.so")


def available(variant: str = "O2") -> bool:
    return os.path.exists(ref_path(variant))


class Ref:
    """One dlopen'ed copy of the reference with its lazy tables pre-warmed."""

    def __init__(self, variant: str = "O2"):
        self.lib = lib = C.CDLL(ref_path(variant))
        fp = C.POINTER(C.c_float)
        lib.fft.argtypes = [fp, fp, fp]
        lib.calculate_bins.argtypes = [fp, fp, fp]
        lib.get_Hann_window.restype = fp
        lib.resample.argtypes = [fp, C.c_uint]
        lib.resample.restype = C.c_void_p
        lib.normalize.argtypes = [fp, C.c_uint]
        lib.build_spectral_images.argtypes = [fp, C.c_uint, C.POINTER(C.POINTER(SpectralImages))]
        lib.apply_Haar_transform.argtypes = [C.POINTER(SpectralImages)]
        lib.transform_image.argtypes = [C.c_void_p]
        lib.build_raw_fingerprints.argtypes = [C.POINTER(SpectralImages)]
        lib.build_raw_fingerprints.restype = C.POINTER(RawFingerprints)
        lib.build_signatures.argtypes = [C.POINTER(RawFingerprints)]
        lib.build_signatures.restype = C.POINTER(Signatures)
        lib.free_signatures.argtypes = [C.POINTER(Signatures)]
        lib.free_rawfingerprints.argtypes = [C.POINTER(RawFingerprints)]
        lib.free_spectral_images.argtypes = [C.POINTER(SpectralImages)]
        lib.get_permutation.argtypes = [C.c_uint]
        lib.get_permutation.restype = C.POINTER(C.c_uint16)
        lib.generate_fingerprint.argtypes = [C.c_char_p, C.POINTER(C.POINTER(Signatures)),
                                             C.POINTER(C.c_char_p), C.POINTER(C.c_char_p), C.POINTER(C.c_char_p)]
        lib.generate_fingerprint_from_samples.argtypes = [fp, C.c_uint, C.POINTER(C.POINTER(Signatures))]
        lib.read_index.argtypes = [C.c_char_p, C.POINTER(C.POINTER(Index))]
        lib.free_index.argtypes = [C.POINTER(Index)]
        lib.create_hash_tables.argtypes = [C.POINTER(Index)]
        lib.create_hash_tables.restype = C.c_void_p
        lib.free_hash_tables.argtypes = [C.c_void_p]
        lib.get_matches.argtypes = [C.c_void_p, C.POINTER(C.c_uint8), C.POINTER(C.POINTER(SignatureList))]
        lib.free_signature_list.argtypes = [C.POINTER(SignatureList)]
        lib.search.argtypes = [C.POINTER(Signatures), C.POINTER(Index), C.c_void_p, C.c_int]
        lib.convert_samples.argtypes = [C.POINTER(C.c_uint8), C.c_uint, C.POINTER(C.POINTER(C.c_float))]
        self.libc = C.CDLL(None)
        self.libc.free.argtypes = [C.c_void_p]
        self._prewarm()

    # -- protocol ---------------------------------------------------------------------
    def _prewarm(self) -> None:
        z = np.zeros(2048, np.float32)
        re = np.zeros(2048, np.float32)
        im = np.zeros(2048, np.float32)
        b = np.zeros(32, np.float32)
        self.lib.fft(_fp(z), _fp(re), _fp(im))
        self.lib.calculate_bins(_fp(re), _fp(im), _fp(b))
        self.lib.get_Hann_window()

    # -- stage functions --------------------------------------------------------------
    def hann(self) -> np.ndarray:
        return np.ctypeslib.as_array(self.lib.get_Hann_window(), (2048,)).copy()

    def permutations(self) -> np.ndarray:
        return np.stack([np.ctypeslib.as_array(self.lib.get_permutation(i), (255,)).copy() for i in range(100)])

    def fft(self, x: np.ndarray):
        x = np.ascontiguousarray(x, np.float32)
        re = np.empty(2048, np.float32)
        im = np.empty(2048, np.float32)
        self.lib.fft(_fp(x), _fp(re), _fp(im))
        return re, im

    def bins(self, re, im) -> np.ndarray:
        b = np.empty(32, np.float32)
        self.lib.calculate_bins(_fp(re), _fp(im), _fp(b))
        return b

    def resample(self, s44: np.ndarray) -> np.ndarray:
        s44 = np.ascontiguousarray(s44, np.float32)
        p = self.lib.resample(_fp(s44), len(s44))
        out = np.ctypeslib.as_array(C.cast(p, C.POINTER(C.c_float)), (len(s44) // 8,)).copy()
        self.libc.free(p)
        return out

    def normalize(self, s: np.ndarray) -> np.ndarray:
        s = np.array(s, np.float32, copy=True)
        self.lib.normalize(_fp(s), len(s))
        return s

    def downmix(self, pcm: np.ndarray) -> np.ndarray:
        """wav.c:358-375 restated with numpy scalars semantics (int sum, f32 div, f64 div)."""
        pcm = np.asarray(pcm, np.int16)
        s = pcm.astype(np.int32).sum(axis=1)
        t = s.astype(np.float32) / np.float32(pcm.shape[1])
        return (t.astype(np.float64) / 32767.0).astype(np.float32)

    def stages_from_samples(self, s5512: np.ndarray, want=("images", "haar", "raw", "sigs")):
        """Run build_spectral_images → Haar → raw fingerprints → signatures on 5512 Hz floats."""
        s = np.ascontiguousarray(s5512, np.float32)
        out = {}
        pimg = C.POINTER(SpectralImages)()
        rc = self.lib.build_spectral_images(_fp(s), len(s), C.byref(pimg))
        out["rc"] = rc
        if rc != 0:
            return out
        n = pimg.contents.n_images
        view = np.ctypeslib.as_array(C.cast(pimg.contents.images, C.POINTER(C.c_float)), (n, IMG_ELEMS))
        if "images" in want:
            out["images"] = view.copy()
        self.lib.apply_Haar_transform(pimg)
        if "haar" in want:
            out["haar"] = view.copy()
        praw = self.lib.build_raw_fingerprints(pimg)
        self.lib.free_spectral_images(pimg)
        raw = np.ctypeslib.as_array(C.cast(praw.contents.fingerprints, C.POINTER(C.c_uint8)), (n, 1025))
        if "raw" in want:
            out["silence"] = raw[:, 0].copy()
            out["raw"] = raw[:, 1:].copy()
        psig = self.lib.build_signatures(praw)
        self.lib.free_rawfingerprints(praw)
        out["sigs"] = _sigs_to_np(psig)
        self.lib.free_signatures(psig)
        return out

    def fingerprint_wav(self, path: str):
        """generate_fingerprint(path) → (rc, signatures[n,100] u8, artist, title, album)."""
        psig = C.POINTER(Signatures)()
        a, t, al = C.c_char_p(), C.c_char_p(), C.c_char_p()
        devnull = os.open(os.devnull, os.O_WRONLY)
        saved = os.dup(2)
        os.dup2(devnull, 2)   # the reference prints progress on stderr
        try:
            rc = self.lib.generate_fingerprint(path.encode(), C.byref(psig), C.byref(a), C.byref(t), C.byref(al))
        finally:
            os.dup2(saved, 2)
            os.close(saved)
            os.close(devnull)
        if rc != 0:
            return rc, None, None, None, None
        sigs = _sigs_to_np(psig)
        self.lib.free_signatures(psig)
        return rc, sigs, a.value, t.value, al.value

    def fingerprint_pcm(self, pcm: np.ndarray):
        """Whole index path on in-memory PCM via the exported stage symbols (no WAV file)."""
        s44 = self.downmix(pcm)
        s = self.normalize(self.resample(s44))
        if len(s) < 2048:
            return -5, None
        r = self.stages_from_samples(s, want=())
        if r["rc"] != 0:
            return r["rc"], None
        return 0, r["sigs"]

    # -- search side ------------------------------------------------------------------
    def load_index(self, path: str):
        pidx = C.POINTER(Index)()
        rc = self.lib.read_index(path.encode(), C.byref(pidx))
        return rc, pidx

    def search(self, sigs: np.ndarray, pidx, lsh) -> int:
        sigs = np.ascontiguousarray(sigs, np.uint8)
        s = Signatures(len(sigs), C.cast(sigs.ctypes.data, C.POINTER(C.c_uint8 * SIG_LEN)))
        return self.lib.search(C.byref(s), pidx, lsh, 0)

    def get_matches(self, lsh, sig: np.ndarray):
        sig = np.ascontiguousarray(sig, np.uint8)
        plist = C.POINTER(SignatureList)()
        n = self.lib.get_matches(lsh, sig.ctypes.data_as(C.POINTER(C.c_uint8)), C.byref(plist))
        out = []
        p = plist
        while p:
            out.append((p.contents.entry_index, p.contents.signature_index))
            p = p.contents.next
        self.lib.free_signature_list(plist)
        assert n == len(out)
        return out


def _fp(a: np.ndarray):
    return a.ctypes.data_as(C.POINTER(C.c_float))


def _sigs_to_np(psig) -> np.ndarray:
    n = psig.contents.n_signatures
    if n == 0:
        return np.zeros((0, SIG_LEN), np.uint8)
    return np.ctypeslib.as_array(C.cast(psig.contents.signatures, C.POINTER(C.c_uint8)), (n, SIG_LEN)).copy()
