"""ctypes wrapper around oracle/libmnemo_oracle.so (our C restatement).  TEST INFRASTRUCTURE ONLY."""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
LIB = os.path.join(HERE, "libmnemo_oracle.so")


class MoDb(C.Structure):
    _fields_ = [("n_entries", C.c_uint32), ("entry_start", C.c_void_p), ("sigs", C.c_void_p)]


class MoScore(C.Structure):
    _fields_ = [("entry", C.c_int32), ("score", C.c_float), ("n_matches", C.c_int32)]


def build() -> None:
    subprocess.check_call(["make", "-s", "-C", HERE, "oracle"])


def _p(a, t=C.c_void_p):
    return a.ctypes.data_as(t)


class Oracle:
    def __init__(self):
        if not os.path.exists(LIB):
            build()
        self.lib = lib = C.CDLL(LIB)
        lib.mo_normalize.restype = C.c_float
        lib.mo_logf_glibc.restype = C.c_float
        lib.mo_logf_glibc.argtypes = [C.c_float, C.c_int]
        lib.mo_permutations.restype = C.POINTER(C.c_uint16)
        lib.mo_n_frames.restype = C.c_uint32
        lib.mo_n_images.restype = C.c_uint32
        lib.mo_n_frames.argtypes = [C.c_uint32]
        lib.mo_n_images.argtypes = [C.c_uint32]
        lib.mo_get_matches.restype = C.c_uint32
        lib.mo_logf_check.restype = C.c_uint64
        lib.mo_logf_check.argtypes = [C.c_uint32, C.c_uint32, C.c_int]

    # tables
    def taps(self):
        h = np.empty(31, np.float32); self.lib.mo_taps(_p(h)); return h

    def hann(self):
        w = np.empty(2048, np.float32); self.lib.mo_hann(_p(w)); return w

    def twiddles(self):
        wr = np.empty(1024, np.float32); wi = np.empty(1024, np.float32)
        self.lib.mo_twiddles(_p(wr), _p(wi)); return wr, wi

    def band_edges(self):
        e = np.empty(33, np.uint16); self.lib.mo_band_edges(_p(e)); return e

    def permutations(self):
        return np.ctypeslib.as_array(self.lib.mo_permutations(), (100, 255)).copy()

    # stages
    def downmix(self, pcm):
        pcm = np.ascontiguousarray(pcm, np.int16)
        out = np.empty(pcm.shape[0], np.float32)
        self.lib.mo_downmix(_p(pcm), C.c_uint64(pcm.shape[0]), C.c_int(pcm.shape[1]), _p(out))
        return out

    def resample(self, s44):
        s44 = np.ascontiguousarray(s44, np.float32)
        out = np.empty(len(s44) // 8, np.float32)
        self.lib.mo_resample(_p(s44), C.c_uint32(len(s44)), _p(out))
        return out

    def normalize(self, s):
        s = np.array(s, np.float32, copy=True)
        rms = self.lib.mo_normalize(_p(s), C.c_uint32(len(s)))
        return s, rms

    def fft(self, x):
        x = np.ascontiguousarray(x, np.float32)
        re = np.empty(2048, np.float32); im = np.empty(2048, np.float32)
        self.lib.mo_fft(_p(x), _p(re), _p(im)); return re, im

    def frame_bins(self, frame):
        frame = np.ascontiguousarray(frame, np.float32)
        b = np.empty(32, np.float32); self.lib.mo_frame_bins(_p(frame), _p(b)); return b

    def scale_image(self, win):
        win = np.ascontiguousarray(win, np.float32).reshape(-1)
        img = np.empty(4096, np.float32); self.lib.mo_scale_image(_p(win), _p(img)); return img

    def haar(self, img):
        img = np.array(img, np.float32, copy=True).reshape(-1); self.lib.mo_haar(_p(img)); return img

    def rawfp(self, himg):
        himg = np.ascontiguousarray(himg, np.float32).reshape(-1)
        bits = np.empty(1024, np.uint8); sil = C.c_int(0)
        self.lib.mo_rawfp(_p(himg), _p(bits), C.byref(sil)); return bits, sil.value

    def minhash(self, bits):
        bits = np.ascontiguousarray(bits, np.uint8)
        sig = np.empty(100, np.uint8)
        ok = self.lib.mo_minhash(_p(bits), _p(sig)); return sig, ok

    def logf(self, x: float, variant: int) -> float:
        return self.lib.mo_logf_glibc(C.c_float(x), variant)

    def logf_check(self, lo_bits: int, hi_bits: int, variant: int) -> int:
        return int(self.lib.mo_logf_check(lo_bits, hi_bits, variant))

    def fingerprint_samples(self, s, taps=False):
        s = np.ascontiguousarray(s, np.float32)
        n = len(s)
        ni = max(0, int(self.lib.mo_n_images(self.lib.mo_n_frames(n)))) if n >= 2048 else 0
        ni = ni if ni < (1 << 30) else 0
        sigs = np.empty((max(ni, 1), 100), np.uint8)
        ns = C.c_uint32(0)
        if taps:
            images = np.empty((max(ni, 1), 4096), np.float32); haar = np.empty_like(images)
            raw = np.empty((max(ni, 1), 1024), np.uint8); sil = np.empty(max(ni, 1), np.uint8)
            rc = self.lib.mo_fingerprint_samples(_p(s), C.c_uint32(n), _p(sigs), C.byref(ns), _p(images), _p(haar), _p(raw), _p(sil))
            return rc, sigs[:ns.value].copy(), dict(images=images[:ni], haar=haar[:ni], raw=raw[:ni], silence=sil[:ni])
        rc = self.lib.mo_fingerprint_samples(_p(s), C.c_uint32(n), _p(sigs), C.byref(ns), None, None, None, None)
        return rc, sigs[:ns.value].copy()

    def fingerprint_pcm(self, pcm, want_samples=False):
        pcm = np.ascontiguousarray(pcm, np.int16)
        nfr, ch = pcm.shape
        n = nfr // 8
        ni = int(self.lib.mo_n_images(self.lib.mo_n_frames(n))) if n >= 2048 else 0
        ni = ni if 0 < ni < (1 << 30) else 0
        sigs = np.empty((max(ni, 1), 100), np.uint8)
        ns = C.c_uint32(0)
        s = np.empty(max(n, 1), np.float32)
        rms = C.c_float(0)
        rc = self.lib.mo_fingerprint_pcm(_p(pcm), C.c_uint64(nfr), C.c_int(ch), _p(sigs), C.byref(ns), _p(s), C.byref(rms))
        if want_samples:
            return rc, sigs[:ns.value].copy(), s[:n], rms.value
        return rc, sigs[:ns.value].copy()

    # search
    @staticmethod
    def make_db(sig_lists):
        counts = np.array([len(s) for s in sig_lists], np.uint32)
        start = np.zeros(len(sig_lists) + 1, np.uint32)
        np.cumsum(counts, out=start[1:])
        allsigs = np.ascontiguousarray(np.concatenate([np.asarray(s, np.uint8).reshape(-1, 100) for s in sig_lists]) if len(sig_lists) else np.zeros((0, 100), np.uint8))
        db = MoDb(len(sig_lists), start.ctypes.data, allsigs.ctypes.data)
        db._keep = (start, allsigs)
        return db

    def search(self, db, qsigs):
        qsigs = np.ascontiguousarray(qsigs, np.uint8)
        scores = (MoScore * max(db.n_entries, 1))()
        L = C.c_uint64(0); D = C.c_uint64(0)
        best = self.lib.mo_search(C.byref(db), _p(qsigs), C.c_uint32(len(qsigs)), scores, C.byref(L), C.byref(D))
        sc = np.array([(s.score, s.n_matches) for s in scores[:db.n_entries]], dtype=np.float64).reshape(-1, 2)
        return best, sc, L.value, D.value

    def search_shard(self, shard_db, table_size, qsigs):
        """Per-entry (score, n) of one shard (local entry indices) + last-group records [nq, 5]."""
        qsigs = np.ascontiguousarray(qsigs, np.uint8)
        scores = (MoScore * max(shard_db.n_entries, 1))()
        last = np.zeros((max(len(qsigs), 1), 5), np.uint32)
        self.lib.mo_search_shard(C.byref(shard_db), C.c_uint32(table_size), _p(qsigs), C.c_uint32(len(qsigs)), scores, _p(last))
        sc = np.array([(s.score, s.n_matches) for s in scores[:shard_db.n_entries]], dtype=np.float64).reshape(-1, 2)
        return sc, last[:len(qsigs)]

    def select(self, score, n_matches):
        n = len(score)
        arr = (MoScore * max(n, 1))()
        for i in range(n):
            arr[i].entry = i; arr[i].score = float(score[i]); arr[i].n_matches = int(n_matches[i])
        return self.lib.mo_select(arr, C.c_uint32(n))

    def get_matches(self, db, sig, cap=1 << 20):
        sig = np.ascontiguousarray(sig, np.uint8)
        out = np.empty((cap, 2), np.uint32)
        n = self.lib.mo_get_matches(C.byref(db), _p(sig), _p(out), C.c_uint32(cap))
        return [tuple(x) for x in out[:n].tolist()]
