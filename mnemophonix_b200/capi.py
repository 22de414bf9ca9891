"""ctypes binding of libmnemo_cuda.so (include/mnemo_cuda.h).

This is plumbing for tests, bench and the torch.distributed drivers; the product is the
shared library.  There is no CPU fallback: loading fails loudly if the library is not built,
and every call fails if no sm_100 GPU is present.
"""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

PKG = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(PKG, "libmnemo_cuda.so")
SIG_LEN = 100

ERRORS = {0: "SUCCESS", -1: "MEMORY_ERROR", -2: "CANNOT_READ_FILE", -3: "DECODING_ERROR",
          -4: "UNSUPPORTED_WAVE_FORMAT", -5: "FILE_TOO_SMALL", -6: "NOT_A_WAVE_FILE", -7: "NO_MATCH_FOUND"}


class MnemoError(RuntimeError):
    pass


class Track(C.Structure):
    _fields_ = [("pcm", C.c_void_p), ("n_frames", C.c_uint64), ("channels", C.c_uint32), ("reserved", C.c_uint32)]


class Match(C.Structure):
    _fields_ = [("entry", C.c_int32), ("score", C.c_float), ("n_matches", C.c_int32)]


MATCH_DTYPE = np.dtype([("entry", "<i4"), ("score", "<f4"), ("n_matches", "<i4")])


def _matches(out, n: int):
    """ctypes Match array -> [(entry, score, n_matches)] (through numpy: a field access per element costs more than the search)."""
    return np.frombuffer(out, dtype=MATCH_DTYPE, count=n).tolist() if n else []


class Cand(C.Structure):
    _fields_ = [("query", C.c_uint32), ("entry", C.c_int32), ("score", C.c_float), ("n_matches", C.c_int32)]


class LastGroup(C.Structure):
    _fields_ = [("has", C.c_uint32), ("entry", C.c_uint32), ("sig", C.c_uint32), ("hits", C.c_uint32),
                ("score", C.c_uint32)]


_lib = None


def lib() -> C.CDLL:
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise MnemoError(f"{LIB_PATH} is not built (python -m mnemophonix_b200.build); there is no CPU fallback")
        L = C.CDLL(LIB_PATH)
        vp, u32, u64, i32 = C.c_void_p, C.c_uint32, C.c_uint64, C.c_int
        L.mnemo_create.argtypes = [i32, vp, C.POINTER(vp)]
        L.mnemo_create_ex.argtypes = [i32, vp, u32, C.POINTER(vp)]
        L.mnemo_spectra_mode.argtypes = [vp]
        L.mnemo_destroy.argtypes = [vp]
        L.mnemo_host_alloc.argtypes = [vp, C.POINTER(vp), u64]
        L.mnemo_host_free.argtypes = [vp, vp]
        L.mnemo_host_free.restype = None
        L.mnemo_last_error.argtypes = [vp]
        L.mnemo_last_error.restype = C.c_char_p
        L.mnemo_logf_variant.argtypes = [vp, C.POINTER(u32), C.POINTER(u32)]
        L.mnemo_samples_5512.argtypes = [u64]
        L.mnemo_samples_5512.restype = u32
        L.mnemo_frames.argtypes = [u32]
        L.mnemo_frames.restype = u32
        L.mnemo_images.argtypes = [u32]
        L.mnemo_images.restype = u32
        L.mnemo_batch_create.argtypes = [vp, C.POINTER(Track), u32, C.POINTER(vp)]
        L.mnemo_batch_destroy.argtypes = [vp]
        L.mnemo_batch_upload.argtypes = [vp, C.POINTER(Track)]
        L.mnemo_batch_run.argtypes = [vp]
        L.mnemo_batch_sync.argtypes = [vp]
        L.mnemo_batch_max_signatures.argtypes = [vp]
        L.mnemo_batch_max_signatures.restype = u64
        L.mnemo_batch_kernel_launches.argtypes = [vp]
        L.mnemo_batch_kernel_launches.restype = u32
        L.mnemo_batch_download.argtypes = [vp, vp, u64, vp, vp]
        L.mnemo_batch_read.argtypes = [vp, i32, u32, vp, u64]
        L.mnemo_batch_read.restype = C.c_int64
        L.mnemo_batch_enable_taps.argtypes = [vp, i32]
        L.mnemo_batch_set_timing.argtypes = [vp, i32]
        L.mnemo_batch_get_timing.argtypes = [vp, vp, C.POINTER(u32)]
        L.mnemo_index_pcm_batch.argtypes = [vp, C.POINTER(Track), u32, vp, u64, vp, vp]
        L.mnemo_index_samples.argtypes = [vp, vp, u32, vp, u64, C.POINTER(u32)]
        L.mnemo_index_samples_graph.argtypes = [vp]
        L.mnemo_selftest_logf.argtypes = [vp, u32, u32]
        L.mnemo_selftest_logf.restype = C.c_int64
        L.mnemo_selftest_div.argtypes = [vp, i32, u32, u64]
        L.mnemo_selftest_div.restype = C.c_int64
        L.mnemo_selftest_select.argtypes = [vp, vp, vp, u32, u32, u32, vp]
        L.mnemo_tree_cuts.argtypes = [u32, u32, vp, vp, u32]
        L.mnemo_tree_cuts.restype = u32
        L.mnemo_search_dense_begin.argtypes = [vp, vp, vp, u32, vp, vp]
        L.mnemo_search_dense_lists.argtypes = [vp, vp, vp, u32, vp, vp]
        L.mnemo_select_merge.argtypes = [vp, vp, u32, u32, vp]
        if hasattr(L, "mnemo_db_create"):
            L.mnemo_db_create.argtypes = [vp, vp, vp, u32, u32, u32, C.POINTER(vp)]
            L.mnemo_db_destroy.argtypes = [vp]
            L.mnemo_db_table_size.argtypes = [vp]
            L.mnemo_db_table_size.restype = u32
            L.mnemo_db_has_pair_index.argtypes = [vp]
            L.mnemo_db_get_matches.argtypes = [vp, vp, vp, u64]
            L.mnemo_db_get_matches.restype = C.c_int64
            L.mnemo_search_batch.argtypes = [vp, vp, vp, u32, vp, vp]
            L.mnemo_search_shard.argtypes = [vp, vp, vp, u32, vp, u64, C.POINTER(u64), vp, vp]
            L.mnemo_search_finish.argtypes = [vp, u64, vp, u32, vp, u32, u32, vp]
            L.mnemo_search_scores.argtypes = [vp, vp, vp, u32, vp, vp]
            L.mnemo_db_tile_queries.argtypes = [vp]
            L.mnemo_db_tile_queries.restype = u32
            L.mnemo_search_dense_begin_dev.argtypes = [vp, vp, vp, u32, vp]
            L.mnemo_search_dense_lists_dev.argtypes = [vp, vp, u32, u32, vp, u32, vp, vp]
            L.mnemo_shard_record_bytes.argtypes = [u32, u32]
            L.mnemo_shard_record_bytes.restype = u64
            L.mnemo_select_merge_dev.argtypes = [vp, vp, u32, u32, u32, u32, vp]
            L.mnemo_search_dense_stats.argtypes = [vp, vp]
            L.mnemo_search_dense_read.argtypes = [vp, u32, vp, vp]
        _lib = L
    return _lib


def _tracks(pcms) -> tuple[C.Array, list]:
    arr = (Track * max(len(pcms), 1))()
    keep = []
    for i, p in enumerate(pcms):
        if isinstance(p, np.ndarray):
            if p.dtype != np.int16 or p.ndim != 2 or not p.flags.c_contiguous:
                raise ValueError("pcm must be C-contiguous int16 [n_frames, channels]")
            keep.append(p)
            arr[i] = Track(p.ctypes.data, p.shape[0], p.shape[1], 0)
        else:  # (ptr, n_frames, channels): e.g. a pinned torch tensor's data_ptr()
            ptr, n, ch = p
            arr[i] = Track(ptr, n, ch, 0)
    return arr, keep


class Context:
    """One mnemo_ctx.  stream: raw cudaStream_t handle (int) or None."""

    def __init__(self, device: int = 0, stream: int | None = None, fast_spectra: bool = False):
        """fast_spectra: the tolerance mode of the frame -> band-energy kernel (MNEMO_FAST_SPECTRA in mnemo_cuda.h)."""
        self.h = C.c_void_p()
        if fast_spectra:
            rc = lib().mnemo_create_ex(device, C.c_void_p(stream) if stream else None, 1, C.byref(self.h))
        else:
            rc = lib().mnemo_create(device, C.c_void_p(stream) if stream else None, C.byref(self.h))
        if rc != 0:
            raise MnemoError(f"mnemo_create(device={device}) failed: {ERRORS.get(rc, rc)} (no GPU? no CPU fallback exists)")
        self.device = device

    def close(self):
        if self.h:
            lib().mnemo_destroy(self.h)
            self.h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _check(self, rc, what):
        if rc != 0:
            raise MnemoError(f"{what}: {ERRORS.get(rc, rc)}: {lib().mnemo_last_error(self.h).decode()}")

    def pinned(self, nbytes: int) -> np.ndarray:
        """A page-locked host buffer (mnemo_host_alloc) as a uint8 array; freed when the array and its views are gone
        (while this context is still open).  Copies from it run at the full host -> device rate, a pageable source is
        staged by the driver at a third of that."""
        import weakref
        ptr = C.c_void_p()
        self._check(lib().mnemo_host_alloc(self.h, C.byref(ptr), max(int(nbytes), 1)), "mnemo_host_alloc")
        raw = (C.c_uint8 * max(int(nbytes), 1)).from_address(ptr.value)
        ctx, addr = self, ptr.value

        def release():
            if ctx.h:
                lib().mnemo_host_free(ctx.h, C.c_void_p(addr))
        weakref.finalize(raw, release)
        return np.ctypeslib.as_array(raw)[:nbytes]

    def logf_variant(self):
        p, m = C.c_uint32(), C.c_uint32()
        v = lib().mnemo_logf_variant(self.h, C.byref(p), C.byref(m))
        return v, p.value, m.value

    def selftest_logf(self, lo_bits: int, hi_bits: int) -> int:
        return lib().mnemo_selftest_logf(self.h, lo_bits, hi_bits)

    def selftest_div(self, which: int, lo_bits: int, n: int) -> int:
        return lib().mnemo_selftest_div(self.h, which, lo_bits, n)

    def selftest_select(self, score: np.ndarray, n_matches: np.ndarray, entry_base: int = 0):
        """The GPU ranking kernel on dense per-entry rows [n_queries, n_entries] (u32): list of (entry, score, n)."""
        score = np.ascontiguousarray(score, np.uint32)
        n_matches = np.ascontiguousarray(n_matches, np.uint32)
        nq, ne = score.shape
        out = (Match * max(nq, 1))()
        self._check(lib().mnemo_selftest_select(self.h, score.ctypes.data, n_matches.ctypes.data, nq, ne, entry_base, out),
                    "mnemo_selftest_select")
        return _matches(out, nq)

    # -- indexing -------------------------------------------------------------------
    def index_pcm_batch(self, pcms):
        """Host buffers in, per-track signature arrays out (list of [n,100] u8) + status list."""
        arr, keep = _tracks(pcms)
        cap = sum(int(lib().mnemo_images(lib().mnemo_samples_5512(arr[i].n_frames))) for i in range(len(pcms)))
        sigs = np.empty((max(cap, 1), SIG_LEN), np.uint8)
        n = np.zeros(max(len(pcms), 1), np.uint32)
        st = np.zeros(max(len(pcms), 1), np.int32)
        rc = lib().mnemo_index_pcm_batch(self.h, arr, len(pcms), sigs.ctypes.data, cap, n.ctypes.data, st.ctypes.data)
        self._check(rc, "mnemo_index_pcm_batch")
        return _split(sigs, n[:len(pcms)]), st[:len(pcms)].tolist()

    def index_samples(self, samples: np.ndarray):
        s = np.ascontiguousarray(samples, np.float32)
        cap = int(lib().mnemo_images(len(s)))
        sigs = np.empty((max(cap, 1), SIG_LEN), np.uint8)
        n = C.c_uint32(0)
        rc = lib().mnemo_index_samples(self.h, s.ctypes.data, len(s), sigs.ctypes.data, cap, C.byref(n))
        if rc not in (0, -5):
            self._check(rc, "mnemo_index_samples")
        return rc, sigs[:n.value].copy()

    @property
    def live_graph(self) -> bool:
        """True if the last index_samples length runs as one captured CUDA graph."""
        return bool(lib().mnemo_index_samples_graph(self.h))

    def batch(self, pcms) -> "Batch":
        return Batch(self, pcms)


def _split(sigs: np.ndarray, counts) -> list:
    out, o = [], 0
    for c in counts:
        out.append(sigs[o:o + int(c)].copy())
        o += int(c)
    return out


class Batch:
    READ_DTYPES = {0: np.float32, 1: np.float32, 2: np.float32, 3: np.uint8, 4: np.uint8, 5: np.uint8,
                   6: np.float32, 7: np.float32, 8: np.float32, 9: np.uint8}

    def __init__(self, ctx: Context, pcms):
        self.ctx = ctx
        self.arr, self._keep = _tracks(pcms)
        self.n = len(pcms)
        self.h = C.c_void_p()
        ctx._check(lib().mnemo_batch_create(ctx.h, self.arr, self.n, C.byref(self.h)), "mnemo_batch_create")
        self.cap = int(lib().mnemo_batch_max_signatures(self.h))

    def close(self):
        if self.h:
            lib().mnemo_batch_destroy(self.h)
            self.h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def upload(self, pcms=None):
        if pcms is not None:
            self.arr, self._keep = _tracks(pcms)
        self.ctx._check(lib().mnemo_batch_upload(self.h, self.arr), "mnemo_batch_upload")

    def run(self):
        self.ctx._check(lib().mnemo_batch_run(self.h), "mnemo_batch_run")

    def sync(self):
        self.ctx._check(lib().mnemo_batch_sync(self.h), "mnemo_batch_sync")

    def launches(self) -> int:
        return int(lib().mnemo_batch_kernel_launches(self.h))

    def enable_taps(self, on=True):
        self.ctx._check(lib().mnemo_batch_enable_taps(self.h, 1 if on else 0), "mnemo_batch_enable_taps")

    STAGES = ("k0_resample", "sumsq_rms", "k1_spectral", "k2_fingerprint", "compact")

    def set_timing(self, on=True):
        self.ctx._check(lib().mnemo_batch_set_timing(self.h, 1 if on else 0), "mnemo_batch_set_timing")

    def get_timing(self) -> tuple[dict, int]:
        ms = (C.c_double * 5)()
        n = C.c_uint32(0)
        self.ctx._check(lib().mnemo_batch_get_timing(self.h, ms, C.byref(n)), "mnemo_batch_get_timing")
        return dict(zip(self.STAGES, list(ms))), n.value

    def download(self, out: np.ndarray | None = None, out_ptr: int | None = None):
        n = np.zeros(max(self.n, 1), np.uint32)
        st = np.zeros(max(self.n, 1), np.int32)
        if out_ptr is None:
            if out is None:
                out = np.empty((max(self.cap, 1), SIG_LEN), np.uint8)
            out_ptr = out.ctypes.data
        rc = lib().mnemo_batch_download(self.h, out_ptr, self.cap, n.ctypes.data, st.ctypes.data)
        self.ctx._check(rc, "mnemo_batch_download")
        return out, n[:self.n], st[:self.n]

    def signatures(self):
        out, n, st = self.download()
        return _split(out, n), st.tolist()

    def record_slots_in_use(self, track: int = 0) -> float:
        """Mean number of group-record slots in use per 8-frame group (1 .. 16): distinct window maxima among the images
        that contain the group (k15_groups.cu).  Image j reads slot[j][p] of group j + p."""
        slot = self.read(9, track).reshape(-1, 16).astype(np.int64)
        ni = slot.shape[0]
        if ni == 0:
            return 0.0
        used = np.zeros(ni + 15, np.int64)
        for p in range(16):
            np.maximum(used[p:p + ni], slot[:, p] + 1, out=used[p:p + ni])
        return float(used.mean())

    def read(self, what: int, track: int) -> np.ndarray:
        d = self.arr[track]
        n5 = int(lib().mnemo_samples_5512(d.n_frames))
        nf = int(lib().mnemo_frames(n5))
        ni = int(lib().mnemo_images(n5))
        count = {0: n5, 1: 1, 2: nf * 32, 3: ni, 4: ni * 100, 5: ni * 1024, 6: ni * 4096, 7: ni * 4096, 8: 1, 9: ni * 16}[what]
        buf = np.empty(max(count, 1), self.READ_DTYPES[what])
        got = lib().mnemo_batch_read(self.h, what, track, buf.ctypes.data, buf.nbytes)
        if got < 0:
            raise MnemoError(f"mnemo_batch_read({what}) failed: {got}: {lib().mnemo_last_error(self.ctx.h).decode()}")
        return buf[:count]


# ---------------------------------------------------------------------------- search ----

def _concat(sig_lists):
    counts = np.array([len(s) for s in sig_lists], np.uint32)
    start = np.zeros(len(sig_lists) + 1, np.uint32)
    np.cumsum(counts, out=start[1:])
    if len(sig_lists):
        allsigs = np.ascontiguousarray(np.concatenate([np.asarray(s, np.uint8).reshape(-1, SIG_LEN) for s in sig_lists]))
    else:
        allsigs = np.zeros((0, SIG_LEN), np.uint8)
    if allsigs.shape[0] == 0:
        allsigs = np.zeros((1, SIG_LEN), np.uint8)[:0]
    return allsigs, start


class Db:
    """LSH database (one shard) resident on the GPU: create_hash_tables (lsh.c:55-86)."""

    def __init__(self, ctx: Context, entry_sigs, table_size: int = 0, entry_base: int = 0):
        self.ctx = ctx
        sigs, start = _concat(entry_sigs)
        self.n_entries = len(entry_sigs)
        self.entry_base = entry_base
        self.h = C.c_void_p()
        buf = sigs if sigs.shape[0] else np.zeros((1, SIG_LEN), np.uint8)
        rc = lib().mnemo_db_create(ctx.h, buf.ctypes.data, start.ctypes.data, self.n_entries, table_size, entry_base,
                                   C.byref(self.h))
        ctx._check(rc, "mnemo_db_create")

    def close(self):
        if self.h:
            lib().mnemo_db_destroy(self.h)
            self.h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    @property
    def table_size(self) -> int:
        return int(lib().mnemo_db_table_size(self.h))

    @property
    def has_pair_index(self) -> bool:
        return bool(lib().mnemo_db_has_pair_index(self.h))

    def get_matches(self, sig: np.ndarray, cap: int = 1 << 20):
        sig = np.ascontiguousarray(sig, np.uint8)
        out = np.empty((cap, 2), np.uint32)
        n = lib().mnemo_db_get_matches(self.h, sig.ctypes.data, out.ctypes.data, cap)
        if n < 0:
            self.ctx._check(int(n), "mnemo_db_get_matches")
        return [tuple(x) for x in out[:min(n, cap)].tolist()]

    @staticmethod
    def prepare(queries, ctx: "Context | None" = None):
        """Concatenate a list of [n,100] u8 arrays once; the result can be passed to search/search_shard.
        ctx: put the concatenated signatures into page-locked memory of that context (Context.pinned)."""
        sigs, start = _concat(queries)
        if ctx is not None and sigs.shape[0]:
            buf = ctx.pinned(sigs.size).reshape(sigs.shape)
            buf[...] = sigs
            sigs = buf
        return sigs, start

    def search(self, queries):
        """queries: list of [n,100] u8 arrays (or Db.prepare(list)).  Returns (results list of
        (entry, score, n_matches), (L, D))."""
        qs, qstart = queries if isinstance(queries, tuple) else _concat(queries)
        nq = len(qstart) - 1
        out = (Match * max(nq, 1))()
        stats = (C.c_uint64 * 2)()
        buf = qs if qs.shape[0] else np.zeros((1, SIG_LEN), np.uint8)
        rc = lib().mnemo_search_batch(self.h, buf.ctypes.data, qstart.ctypes.data, nq, out, stats)
        self.ctx._check(rc, "mnemo_search_batch")
        return _matches(out, nq), (stats[0], stats[1])

    def scores(self, queries):
        """Dense per-entry rows of search.c:55-59 for a (small) batch: (score[nq, n_entries], n_matches[nq, n_entries]) u32."""
        qs, qstart = queries if isinstance(queries, tuple) else _concat(queries)
        nq = len(qstart) - 1
        sc = np.zeros((max(nq, 1), max(self.n_entries, 1)), np.uint32)
        nm = np.zeros_like(sc)
        buf = qs if qs.shape[0] else np.zeros((1, SIG_LEN), np.uint8)
        rc = lib().mnemo_search_scores(self.h, buf.ctypes.data, qstart.ctypes.data, nq, sc.ctypes.data, nm.ctypes.data)
        self.ctx._check(rc, "mnemo_search_scores")
        return sc[:nq, :self.n_entries], nm[:nq, :self.n_entries]

    @property
    def tile_queries(self) -> int:
        return int(lib().mnemo_db_tile_queries(self.h))

    # -- sharded search ranked on the GPUs (DESIGN.md "search, multi-GPU") --
    def dense_begin_dev(self, queries, d_has_ptr: int):
        """Device-resident form: probe this shard, has flags into device memory at d_has_ptr (async on the context's stream)."""
        qs, qstart = queries if isinstance(queries, tuple) else _concat(queries)
        nq = len(qstart) - 1
        buf = qs if qs.shape[0] else np.zeros((1, SIG_LEN), np.uint8)
        rc = lib().mnemo_search_dense_begin_dev(self.h, buf.ctypes.data, qstart.ctypes.data, nq, C.c_void_p(d_has_ptr))
        self.ctx._check(rc, "mnemo_search_dense_begin_dev")
        self._dense_nq = nq

    def dense_lists_dev(self, d_has_all_ptr: int, n_shards: int, shard: int, node_bounds, d_lists_ptr: int, d_counts_ptr: int):
        node_bounds = np.ascontiguousarray(node_bounds, np.uint32)
        rc = lib().mnemo_search_dense_lists_dev(self.h, C.c_void_p(d_has_all_ptr), n_shards, shard, node_bounds.ctypes.data,
                                                len(node_bounds) - 1, C.c_void_p(d_lists_ptr), C.c_void_p(d_counts_ptr))
        self.ctx._check(rc, "mnemo_search_dense_lists_dev")

    def dense_stats(self):
        stats = (C.c_uint64 * 2)()
        self.ctx._check(lib().mnemo_search_dense_stats(self.h, stats), "mnemo_search_dense_stats")
        return int(stats[0]), int(stats[1])

    def dense_read(self, query: int):
        """Accumulator row of one query of the held batch: (score[n_entries], n_matches[n_entries]) u32, this shard's entries."""
        sc = np.zeros(max(self.n_entries, 1), np.uint32)
        nm = np.zeros_like(sc)
        self.ctx._check(lib().mnemo_search_dense_read(self.h, query, sc.ctypes.data, nm.ctypes.data), "mnemo_search_dense_read")
        return sc[:self.n_entries], nm[:self.n_entries]

    def dense_begin(self, queries):
        """Probe this shard; the accumulators stay on the device.  Returns (has[n_query_sigs] u8, (L, D))."""
        qs, qstart = queries if isinstance(queries, tuple) else _concat(queries)
        nq, nqs = len(qstart) - 1, int(qstart[-1])
        has = np.zeros(max(nqs, 1), np.uint8)
        stats = (C.c_uint64 * 2)()
        buf = qs if qs.shape[0] else np.zeros((1, SIG_LEN), np.uint8)
        rc = lib().mnemo_search_dense_begin(self.h, buf.ctypes.data, qstart.ctypes.data, nq, has.ctypes.data, stats)
        self.ctx._check(rc, "mnemo_search_dense_begin")
        self._dense_nq = nq
        return has[:nqs], (stats[0], stats[1])

    def dense_lists(self, has_higher, node_bounds):
        """Fix-up + ranking of this shard's nodes.  Returns (lists [n_nodes, nq, 10] SEL_DTYPE, counts [n_nodes, nq])."""
        node_bounds = np.ascontiguousarray(node_bounds, np.uint32)
        n_nodes, nq = len(node_bounds) - 1, self._dense_nq
        lists = np.zeros((n_nodes, max(nq, 1), 10), SEL_DTYPE)
        counts = np.zeros((n_nodes, max(nq, 1)), np.uint32)
        hh = None if has_higher is None else np.ascontiguousarray(has_higher, np.uint8)
        rc = lib().mnemo_search_dense_lists(self.h, None if hh is None else hh.ctypes.data, node_bounds.ctypes.data, n_nodes,
                                            lists.ctypes.data, counts.ctypes.data)
        self.ctx._check(rc, "mnemo_search_dense_lists")
        return lists[:, :nq], counts[:, :nq]

    def search_shard(self, queries, cap: int | None = None):
        """Per-shard candidate records + last-group records, as numpy structured arrays."""
        qs, qstart = queries if isinstance(queries, tuple) else _concat(queries)
        nq = len(qstart) - 1
        nqs = int(qstart[-1])
        cap = cap or max(1024, nq * 64)
        while True:
            cands = (Cand * cap)()
            last = (LastGroup * max(nqs, 1))()
            n = C.c_uint64(0)
            stats = (C.c_uint64 * 2)()
            buf = qs if qs.shape[0] else np.zeros((1, SIG_LEN), np.uint8)
            rc = lib().mnemo_search_shard(self.h, buf.ctypes.data, qstart.ctypes.data, nq, cands, cap, C.byref(n), last, stats)
            if rc == -1 and n.value > cap:
                cap = int(n.value)
                continue
            self.ctx._check(rc, "mnemo_search_shard")
            break
        c = np.frombuffer(cands, dtype=CAND_DTYPE, count=int(n.value)).copy()
        l = np.frombuffer(last, dtype=LAST_DTYPE, count=nqs).copy()
        return c, l, qstart, (stats[0], stats[1])


SEL_DTYPE = np.dtype([("entry", "<i4"), ("score", "<f4"), ("n_matches", "<i4"), ("avg", "<f4")])


def tree_cuts(n_entries_total: int, n_shards: int):
    """Entry ranges of shards that sit on node boundaries of the ranking tree: (cuts[n_shards+1], node_lo[n_nodes+1])."""
    n_nodes = 1
    while n_nodes < n_shards:
        n_nodes *= 2
    cuts = np.zeros(n_shards + 1, np.uint32)
    node_lo = np.zeros(n_nodes + 1, np.uint32)
    got = lib().mnemo_tree_cuts(n_entries_total, n_shards, cuts.ctypes.data, node_lo.ctypes.data, n_nodes + 1)
    assert got == n_nodes
    return cuts, node_lo


def shard_record_bytes(nodes_per_shard: int, n_queries: int) -> int:
    return int(lib().mnemo_shard_record_bytes(nodes_per_shard, n_queries))


def select_merge_dev(ctx: Context, d_gathered_ptr: int, n_shards: int, n_nodes: int, nodes_per_shard: int, n_queries: int,
                     as_array: bool = False):
    """Top levels of the ranking tree on the GPU over the all-gathered shard blocks (device memory).
    as_array: the answers as one MATCH_DTYPE array (the bytes the library wrote) instead of a list of tuples."""
    out = (Match * max(n_queries, 1))()
    rc = lib().mnemo_select_merge_dev(ctx.h, C.c_void_p(d_gathered_ptr), n_shards, n_nodes, nodes_per_shard, n_queries, out)
    ctx._check(rc, "mnemo_select_merge_dev")
    if as_array:
        return np.frombuffer(out, dtype=MATCH_DTYPE, count=n_queries).copy() if n_queries else np.zeros(0, MATCH_DTYPE)
    return _matches(out, n_queries)


def select_merge(lists: np.ndarray, counts: np.ndarray):
    """lists [n_nodes, nq, 10] SEL_DTYPE of ALL nodes in node order, counts [n_nodes, nq] -> [(entry, score, n)] per query."""
    lists = np.ascontiguousarray(lists, SEL_DTYPE)
    counts = np.ascontiguousarray(counts, np.uint32)
    n_nodes, nq = counts.shape
    out = (Match * max(nq, 1))()
    rc = lib().mnemo_select_merge(lists.ctypes.data, counts.ctypes.data, n_nodes, nq, out)
    if rc != 0:
        raise MnemoError(f"mnemo_select_merge: {ERRORS.get(rc, rc)}")
    return _matches(out, nq)


CAND_DTYPE = np.dtype([("query", "<u4"), ("entry", "<i4"), ("score", "<f4"), ("n_matches", "<i4")])
LAST_DTYPE = np.dtype([("has", "<u4"), ("entry", "<u4"), ("sig", "<u4"), ("hits", "<u4"), ("score", "<u4")])


def search_finish(cands: np.ndarray, last_all: np.ndarray | None, n_shards: int, qstart: np.ndarray, n_entries_total: int):
    """Host-only merge + exact selection (search.c:170-193).  cands: CAND_DTYPE records of all shards
    concatenated; last_all: LAST_DTYPE [n_shards, n_query_sigs] or None."""
    cands = np.ascontiguousarray(cands, dtype=CAND_DTYPE)
    nq = len(qstart) - 1
    out = (Match * max(nq, 1))()
    qstart = np.ascontiguousarray(qstart, np.uint32)
    lp = None
    if last_all is not None:
        last_all = np.ascontiguousarray(last_all, dtype=LAST_DTYPE)
        lp = last_all.ctypes.data
    rc = lib().mnemo_search_finish(cands.ctypes.data if len(cands) else None, len(cands), lp, n_shards,
                                   qstart.ctypes.data, nq, n_entries_total, out)
    if rc != 0:
        raise MnemoError(f"mnemo_search_finish: {ERRORS.get(rc, rc)}")
    return _matches(out, nq)
