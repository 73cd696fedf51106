"""Scanner: a thin, typed wrapper over one ntl_ctx (one CUDA device, one NanoTel parameter set).

It adds nothing to the C ABI except numpy views; all arithmetic happens in libnanotel_b200.so on the GPU.
"""
from __future__ import annotations

import ctypes as C
from typing import List, Optional, Sequence

import numpy as np

from . import _lib


class NanoTelError(RuntimeError):
    def __init__(self, code: int, message: str):
        super().__init__("libnanotel_b200 error %d: %s" % (code, message))
        self.code = code


class Scanner:
    """One GPU context for a fixed (--patterns, --tvr_patterns, --min_density, --subseq_length, --rc,
    --use_filter, --check_right_edge) set -- the arguments NanoTel.R fixes for a whole run (NanoTel.R:2322-2341)."""

    def __init__(self, patterns, tvr_patterns=None, min_density: float = 0.6, subseq_length: int = 100,
                 rc: bool = False, use_filter: bool = False, right_edge: bool = False, device: int = 0,
                 jit: Optional[bool] = None, debug_stages: bool = False, host_threads: int = 0,
                 devices: Optional[Sequence[int]] = None):
        """jit: None = the specialised span scan kernel if one can be had (precompiled, cached or NVRTC), else the
        generic kernel with a warning; True = fail without it; False = the generic kernel.
        devices: CUDA ordinals to shard every batch over (contiguous shards balanced by bases, gathered in input
        order); None = the single `device`."""
        self._L = _lib.load()
        options = 0
        if jit is False:
            options |= _lib.OPT_NO_JIT
        elif jit is True:
            options |= _lib.OPT_REQUIRE_JIT
        if debug_stages:
            options |= _lib.OPT_DEBUG_STAGES
        self.params = _lib.make_params(patterns, tvr_patterns, min_density, subseq_length, rc, use_filter,
                                       right_edge, device, options, host_threads, devices)
        self.n_tracks = 3 if self.params.n_tvr > 0 else 2
        self.debug_stages = debug_stages
        self._h = C.c_void_p()
        rc_ = self._L.ntl_create(C.byref(self._h), C.byref(self.params))
        if rc_ != _lib.NTL_OK:
            msg = self._L.ntl_last_error(None).decode(errors="replace")
            self._h = C.c_void_p()
            raise NanoTelError(rc_, msg)
        self._n = 0
        self._keep = None

    # -- lifecycle
    def close(self) -> None:
        if getattr(self, "_h", None) and self._h.value:
            self._L.ntl_destroy(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def __enter__(self):
        return self

    def __exit__(self, *a):
        self.close()

    def _check(self, rc: int) -> None:
        if rc != _lib.NTL_OK:
            raise NanoTelError(rc, self._L.ntl_last_error(self._h).decode(errors="replace"))

    @property
    def note(self) -> str:
        """Why the context runs the generic scan kernel ("" when it runs the specialised one)."""
        return self._L.ntl_scan_path_note(self._h).decode(errors="replace")

    @property
    def scan_path(self) -> str:
        """"precompiled" | "cached" | "nvrtc" (specialised span scan) or "generic"."""
        return _lib.SCAN_PATH_NAMES.get(self._L.ntl_scan_path(self._h), "?")

    @property
    def n_devices(self) -> int:
        return int(self._L.ntl_device_count(self._h))

    def shards(self):
        """(bounds[n_devices + 1], devices[n_devices]) of the last batch."""
        g = self.n_devices
        b = np.zeros(g + 1, np.int32)
        d = np.zeros(g, np.int32)
        self._L.ntl_get_shards(self._h, b.ctypes.data, d.ctypes.data, g)
        return b, d

    def geometry(self) -> dict:
        v = [C.c_int32() for _ in range(4)]
        self._L.ntl_get_geometry(self._h, *[C.byref(x) for x in v])
        return dict(zip(("block", "blocks_per_window", "words_per_span", "blocks_per_span"), (x.value for x in v)))

    # -- batches
    @staticmethod
    def _marshal(seqs: Sequence[bytes]):
        n = len(seqs)
        arr = (C.c_char_p * max(n, 1))(*seqs)
        lens = np.fromiter((len(s) for s in seqs), dtype=np.int64, count=n)
        return arr, lens

    def _results_view(self, ptr: C.c_void_p, n: int, out=None) -> np.ndarray:
        """The records of the last batch, copied out of the context's pinned buffer (which the next batch reuses):
        into `out` (RESULT_DTYPE, >= n entries, no allocation and no page faults) or into a fresh array.
        out="view" returns a read-only view of that pinned buffer instead (what the C ABI hands out: valid until the
        next batch on this Scanner)."""
        if n == 0:
            return np.zeros(0, _lib.RESULT_DTYPE)
        buf = (C.c_char * (n * 64)).from_address(ptr.value)
        # one flat memcpy (a structured-dtype .copy() is ~10x slower)
        src = np.frombuffer(buf, dtype=np.uint8, count=n * 64)
        if isinstance(out, str):
            if out != "view":
                raise ValueError('out must be an array, None or "view"')
            v = src.view(_lib.RESULT_DTYPE)
            v.flags.writeable = False
            return v
        if out is None:
            return src.copy().view(_lib.RESULT_DTYPE)
        if out.dtype != _lib.RESULT_DTYPE or out.ndim != 1 or out.shape[0] < n or not out.flags.c_contiguous:
            raise ValueError("out must be a contiguous 1-d RESULT_DTYPE array with at least %d entries" % n)
        np.copyto(out[:n].view(np.uint8), src)
        return out[:n]

    def scan(self, seqs: Sequence[bytes], out=None) -> np.ndarray:
        """ntl_scan_batch: host ASCII reads in, one RESULT_DTYPE record per read out (a copy)."""
        arr, lens = self._marshal(seqs)
        self._keep = (arr, lens, seqs)
        res = C.c_void_p()
        self._check(self._L.ntl_scan_batch(self._h, arr, lens.ctypes.data, len(seqs), C.byref(res)))
        self._n = len(seqs)
        return self._results_view(res, self._n, out)

    @staticmethod
    def _marshal_concat(buf: np.ndarray, offsets: np.ndarray):
        buf = np.ascontiguousarray(buf, dtype=np.uint8)
        offsets = np.ascontiguousarray(offsets, dtype=np.int64)
        ptrs = (np.uint64(buf.ctypes.data) + offsets[:-1].astype(np.uint64)).astype(np.uint64)
        lens = np.diff(offsets).astype(np.int64)
        return buf, ptrs, lens

    def scan_concat(self, buf: np.ndarray, offsets: np.ndarray, out=None) -> np.ndarray:
        """ntl_scan_batch_concat: reads given as one ASCII buffer + offsets (read i = buf[offsets[i]:offsets[i+1]])."""
        buf = np.ascontiguousarray(buf, dtype=np.uint8)
        offsets = np.ascontiguousarray(offsets, dtype=np.int64)
        self._keep = (buf, offsets)
        n = len(offsets) - 1
        res = C.c_void_p()
        self._check(self._L.ntl_scan_batch_concat(self._h, buf.ctypes.data, offsets.ctypes.data, n, C.byref(res)))
        self._n = n
        return self._results_view(res, n, out)

    def scan_pool(self, pool: np.ndarray, start: np.ndarray, width: np.ndarray, biostrings_codes: bool = False,
                  out=None) -> np.ndarray:
        """ntl_scan_batch_pool: reads as an XStringSet holds them (pool of bytes + 1-based start + width)."""
        pool = np.ascontiguousarray(pool, dtype=np.uint8)
        start = np.ascontiguousarray(start, dtype=np.int32)
        width = np.ascontiguousarray(width, dtype=np.int32)
        self._keep = (pool, start, width)
        n = len(start)
        res = C.c_void_p()
        self._check(self._L.ntl_scan_batch_pool(self._h, pool.ctypes.data, start.ctypes.data, width.ctypes.data, n,
                                                int(bool(biostrings_codes)), C.byref(res)))
        self._n = n
        return self._results_view(res, n, out)

    def pack_concat(self, buf: np.ndarray, offsets: np.ndarray) -> None:
        buf, ptrs, lens = self._marshal_concat(buf, offsets)
        self._keep = (buf, ptrs, lens)
        self._check(self._L.ntl_batch_pack(self._h, ptrs.ctypes.data, lens.ctypes.data, len(lens)))
        self._n = len(lens)

    def pack(self, seqs: Sequence[bytes]) -> None:
        arr, lens = self._marshal(seqs)
        self._keep = (arr, lens, seqs)
        self._check(self._L.ntl_batch_pack(self._h, arr, lens.ctypes.data, len(seqs)))
        self._n = len(seqs)

    def upload(self) -> None:
        self._check(self._L.ntl_batch_upload(self._h))

    def run(self) -> None:
        self._check(self._L.ntl_batch_run(self._h))

    def enqueue(self) -> None:
        self._check(self._L.ntl_batch_enqueue(self._h))

    def wait(self) -> None:
        self._check(self._L.ntl_batch_wait(self._h))

    def download(self, out=None) -> np.ndarray:
        res = C.c_void_p()
        self._check(self._L.ntl_batch_download(self._h, C.byref(res)))
        return self._results_view(res, self._n, out)

    def timings(self) -> dict:
        t = _lib.Timings()
        self._check(self._L.ntl_get_timings(self._h, C.byref(t)))
        return {k: getattr(t, k) for k, _ in _lib.Timings._fields_}

    @property
    def stream(self) -> int:
        return int(self._L.ntl_stream(self._h) or 0)

    # -- per-window tables (analyze_subtelos' data.frame, NanoTel.R:740-765)
    def windows(self, read_idx: int, track: int, n_win: Optional[int] = None):
        if n_win is None:
            n_win = self._L.ntl_get_windows(self._h, read_idx, track, 0, None, None, None, None)
            if n_win < 0:
                self._check(n_win)
        st = np.zeros(max(n_win, 1), np.int32)
        en = np.zeros(max(n_win, 1), np.int32)
        cov = np.zeros(max(n_win, 1), np.int32)
        den = np.zeros(max(n_win, 1), np.float64)
        r = self._L.ntl_get_windows(self._h, read_idx, track, n_win, st.ctypes.data, en.ctypes.data,
                                    cov.ctypes.data, den.ctypes.data)
        if r < 0:
            self._check(r)
        return st[:n_win], en[:n_win], cov[:n_win], den[:n_win]

    def window_counts(self, track: int, total: Optional[int] = None) -> np.ndarray:
        """ntl_get_window_counts: covered counts of every window of every read of the last batch (uint16, read after
        read, n_win entries each) -- the bulk form of windows() for whole-batch comparisons."""
        if total is None:
            total = self._L.ntl_get_window_counts(self._h, track, None, 0)
            if total < 0:
                self._check(int(total))
        out = np.zeros(max(int(total), 1), np.uint16)
        r = self._L.ntl_get_window_counts(self._h, track, out.ctypes.data, int(total))
        if r < 0:
            self._check(int(r))
        return out[:int(total)]

    def stages(self, read_idx: int, track: int) -> dict:
        s = _lib.Stage()
        self._check(self._L.ntl_get_stages(self._h, read_idx, track, C.byref(s)))
        return {k: getattr(s, k) for k, _ in _lib.Stage._fields_}


def assign_serials(results: np.ndarray, serial_start: int = 1):
    """ntl_assign_serials: (serial[n], row_order[rows], next_serial_start) -- NanoTel.R:2050-2069, 2234-2258."""
    L = _lib.load()
    res = np.ascontiguousarray(results, dtype=_lib.RESULT_DTYPE)
    n = len(res)
    serial = np.zeros(max(n, 1), np.int32)
    order = np.zeros(max(n, 1), np.int32)
    nxt = C.c_int32(serial_start)
    rows = L.ntl_assign_serials(res.ctypes.data, n, serial_start, serial.ctypes.data, order.ctypes.data, C.byref(nxt))
    if rows < 0:
        raise NanoTelError(rows, "ntl_assign_serials")
    return serial[:n], order[:rows], nxt.value


def count_windows(length: int, subseq_length: int) -> int:
    return _lib.load().ntl_count_windows(int(length), int(subseq_length))
