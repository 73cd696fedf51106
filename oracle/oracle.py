"""ctypes binding of oracle/nanotel_oracle.c -- TEST INFRASTRUCTURE, NOT PRODUCT.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may import this
module.  The product package (telomere-analyzer_b200/nanotel_b200) never does.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess
from dataclasses import dataclass
from typing import List, Optional, Sequence

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "_build", "libnanotel_oracle.so")

MAX_PATTERNS = 16
FLAG_REF_ERROR = 1
FLAG_NO_WINDOWS = 2


class Params(C.Structure):
    _fields_ = [
        ("n_patterns", C.c_int32),
        ("patterns", C.c_char_p * MAX_PATTERNS),
        ("n_tvr", C.c_int32),
        ("tvr", C.c_char_p * MAX_PATTERNS),
        ("min_density", C.c_double),
        ("subseq_length", C.c_int32),
        ("right_edge", C.c_int32),
    ]


class Track(C.Structure):
    _fields_ = [
        ("coarse_start", C.c_int32), ("coarse_end", C.c_int32),
        ("acc_start", C.c_int32), ("acc_end", C.c_int32),
        ("edge_start", C.c_int32), ("edge_end", C.c_int32),
        ("start", C.c_int32), ("end", C.c_int32),
        ("acc_density", C.c_double), ("density", C.c_double),
        ("n_ranges", C.c_int32), ("pad", C.c_int32),
    ]


class Read(C.Structure):
    _fields_ = [
        ("keep", C.c_int32), ("flags", C.c_int32), ("n_win", C.c_int32), ("length", C.c_int32),
        ("t", Track * 3),
    ]


READ_DTYPE = np.dtype({
    "names": ["keep", "flags", "n_win", "length", "t"],
    "formats": ["<i4", "<i4", "<i4", "<i4", (np.dtype({
        "names": ["coarse_start", "coarse_end", "acc_start", "acc_end", "edge_start", "edge_end", "start", "end",
                  "acc_density", "density", "n_ranges", "pad"],
        "formats": ["<i4"] * 8 + ["<f8", "<f8", "<i4", "<i4"],
    }), 3)],
})
assert READ_DTYPE.itemsize == C.sizeof(Read)


def build(force: bool = False) -> str:
    """Compile the oracle with gcc (oracle/Makefile)."""
    src = os.path.join(_HERE, "nanotel_oracle.c")
    if force or not os.path.exists(_SO) or os.path.getmtime(_SO) < os.path.getmtime(src):
        subprocess.run(["make", "-C", _HERE, "-s"], check=True)
    return _SO


_lib = None


def lib():
    global _lib
    if _lib is None:
        build()
        L = C.CDLL(_SO)
        L.ntlo_analyze_read.argtypes = [C.POINTER(Params), C.c_char_p, C.c_int32, C.POINTER(Read),
                                        C.c_void_p, C.c_int32]
        L.ntlo_analyze_read.restype = C.c_int
        L.ntlo_split_telo.argtypes = [C.c_int32, C.c_int32, C.c_void_p, C.c_void_p, C.c_int32]
        L.ntlo_split_telo.restype = C.c_int32
        L.ntlo_count_windows.argtypes = [C.c_int32, C.c_int32]
        L.ntlo_count_windows.restype = C.c_int32
        L.ntlo_filter_read.argtypes = [C.POINTER(Params), C.c_char_p, C.c_int32]
        L.ntlo_filter_read.restype = C.c_int
        L.ntlo_revcomp.argtypes = [C.c_char_p, C.c_int32, C.c_char_p]
        L.ntlo_revcomp.restype = None
        L.ntlo_match_pattern.argtypes = [C.c_char_p, C.c_int32, C.c_char_p, C.c_int32, C.c_int32, C.c_void_p,
                                         C.c_int32]
        L.ntlo_match_pattern.restype = C.c_int32
        L.ntlo_sub_density_ranges.argtypes = [C.c_void_p, C.c_void_p, C.c_int32, C.c_int32, C.c_int32, C.c_int32]
        L.ntlo_sub_density_ranges.restype = C.c_double
        L.ntlo_scan_batch.argtypes = [C.POINTER(Params), C.POINTER(C.c_char_p), C.c_void_p, C.c_int32, C.c_int32,
                                      C.c_int32, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int32]
        L.ntlo_scan_batch.restype = C.c_int
        L.ntlo_assign_serials.argtypes = [C.c_void_p, C.c_int32, C.c_int32, C.c_int32, C.c_void_p, C.c_void_p,
                                          C.POINTER(C.c_int32), C.POINTER(C.c_int32)]
        L.ntlo_assign_serials.restype = C.c_int32
        _lib = L
    return _lib


def _split(p) -> List[bytes]:
    if p is None:
        return []
    if isinstance(p, (str, bytes)):
        p = p.split()
    return [x.encode() if isinstance(x, str) else x for x in p]


def make_params(patterns, tvr_patterns=None, min_density: float = 0.6, subseq_length: int = 100,
                right_edge: bool = False) -> Params:
    pats, tvr = _split(patterns), _split(tvr_patterns)
    P = Params()
    P.n_patterns = len(pats)
    for i, x in enumerate(pats):
        P.patterns[i] = x
    P.n_tvr = len(tvr)
    for i, x in enumerate(tvr):
        P.tvr[i] = x
    P.min_density = float(min_density)
    P.subseq_length = int(subseq_length)
    P.right_edge = int(bool(right_edge))
    P._keep = (pats, tvr)
    return P


def count_windows(length: int, S: int) -> int:
    return lib().ntlo_count_windows(length, S)


def split_telo(length: int, S: int):
    n = count_windows(length, S)
    st = np.zeros(max(n, 1), np.int32)
    en = np.zeros(max(n, 1), np.int32)
    lib().ntlo_split_telo(length, S, st.ctypes.data, en.ctypes.data, n)
    return st[:n], en[:n]


def revcomp(seq: bytes) -> bytes:
    out = C.create_string_buffer(len(seq))
    lib().ntlo_revcomp(seq, len(seq), out)
    return out.raw


def match_pattern(seq: bytes, pat: str, max_mismatch: int = 0, fixed: bool = True) -> np.ndarray:
    cap = len(seq) + 4
    st = np.zeros(cap, np.int32)
    n = lib().ntlo_match_pattern(seq, len(seq), pat.encode(), max_mismatch, int(fixed), st.ctypes.data, cap)
    if n < 0:
        raise ValueError("bad pattern or sequence")
    return st[:n].copy()


def sub_density_ranges(starts, ends, L: int, a: int, b: int) -> float:
    """get_sub_density(IRanges(a, b), IRanges(starts, ends)) on a subject of length L (NanoTel.R:449-468)."""
    st = np.ascontiguousarray(starts, np.int32)
    en = np.ascontiguousarray(ends, np.int32)
    return float(lib().ntlo_sub_density_ranges(st.ctypes.data, en.ctypes.data, len(st), L, a, b))


def filter_read(P: Params, seq: bytes) -> bool:
    r = lib().ntlo_filter_read(C.byref(P), seq, len(seq))
    if r < 0:
        raise ValueError("oracle filter error %d" % r)
    return bool(r)


@dataclass
class ReadResult:
    rec: np.void                    # READ_DTYPE scalar
    win_counts: np.ndarray          # [n_tracks, n_win] int32


def analyze_read(P: Params, seq: bytes) -> ReadResult:
    n_tracks = 3 if P.n_tvr > 0 else 2
    n_win = count_windows(len(seq), P.subseq_length)
    wc = np.zeros((n_tracks, max(n_win, 1)), np.int32)
    out = Read()
    rc = lib().ntlo_analyze_read(C.byref(P), seq, len(seq), C.byref(out), wc.ctypes.data, wc.shape[1])
    if rc:
        raise ValueError("oracle error %d" % rc)
    rec = np.frombuffer(bytes(out), dtype=READ_DTYPE)[0]
    return ReadResult(rec, wc[:, :n_win])


def scan_batch(P: Params, seqs: Sequence[bytes], do_rc: bool = False, use_filter: bool = False,
               n_threads: int = 1, want_windows: bool = True):
    """Returns (records[READ_DTYPE n], pass[uint8 n], win_off[int64 n+1], win_counts[int32])."""
    n = len(seqs)
    n_tracks = 3 if P.n_tvr > 0 else 2
    lens = np.array([len(s) for s in seqs], np.int32)
    arr = (C.c_char_p * n)(*seqs)
    out = np.zeros(n, READ_DTYPE)
    passed = np.zeros(n, np.uint8)
    if want_windows:
        nw = np.array([count_windows(int(l), P.subseq_length) for l in lens], np.int64)
        win_off = np.zeros(n + 1, np.int64)
        np.cumsum(nw * n_tracks, out=win_off[1:])
        wc = np.zeros(max(int(win_off[-1]), 1), np.int32)
        wo_p, wc_p = win_off.ctypes.data, wc.ctypes.data
    else:
        win_off, wc, wo_p, wc_p = None, None, None, None
    rc = lib().ntlo_scan_batch(C.byref(P), arr, lens.ctypes.data, n, int(do_rc), int(use_filter),
                               out.ctypes.data, passed.ctypes.data, wo_p, wc_p, int(n_threads))
    if rc:
        raise ValueError("oracle batch error %d" % rc)
    return out, passed, win_off, wc


def assign_serials(keep: np.ndarray, serial_start: int, prev_max_serial: int = 0):
    keep = np.ascontiguousarray(keep, np.int32)
    n = len(keep)
    serial = np.zeros(max(n, 1), np.int32)
    order = np.zeros(max(n, 1), np.int32)
    nxt, mx = C.c_int32(), C.c_int32()
    rows = lib().ntlo_assign_serials(keep.ctypes.data, n, serial_start, prev_max_serial, serial.ctypes.data,
                                     order.ctypes.data, C.byref(nxt), C.byref(mx))
    return serial[:n], order[:rows], nxt.value, mx.value
