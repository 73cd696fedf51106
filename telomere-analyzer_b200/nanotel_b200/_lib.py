"""ctypes binding of libnanotel_b200.so (include/nanotel_b200.h).

The shared library is the product; this module only declares its C ABI for the Python host mirror.  There is no
CPU fallback anywhere in this package: if the library is missing the import fails loudly, and if no CUDA device is
present `ntl_create` fails with NTL_ERR_CUDA.
"""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libnanotel_b200.so")

NTL_OK = 0
NTL_ERR_ARG, NTL_ERR_PATTERN, NTL_ERR_SEQUENCE, NTL_ERR_CUDA, NTL_ERR_NOMEM, NTL_ERR_JIT, NTL_ERR_STATE = range(-1, -8, -1)

OPT_NO_JIT, OPT_REQUIRE_JIT, OPT_DEBUG_STAGES = 1, 2, 4
SCAN_GENERIC, SCAN_PRECOMPILED, SCAN_CACHED, SCAN_NVRTC = 0, 1, 2, 3
SCAN_PATH_NAMES = {0: "generic", 1: "precompiled", 2: "cached", 3: "nvrtc"}
READ_KEEP, READ_FILTERED, READ_REF_ERROR, READ_NO_WINDOWS, READ_IUPAC = 1, 2, 4, 8, 16


class Params(C.Structure):
    _fields_ = [
        ("n_patterns", C.c_int32), ("patterns", C.POINTER(C.c_char_p)),
        ("n_tvr", C.c_int32), ("tvr_patterns", C.POINTER(C.c_char_p)),
        ("min_density", C.c_double), ("subseq_length", C.c_int32),
        ("rc", C.c_int32), ("use_filter", C.c_int32), ("right_edge", C.c_int32),
        ("device", C.c_int32), ("options", C.c_uint32), ("host_threads", C.c_int32), ("n_devices", C.c_int32),
        ("device_ids", C.POINTER(C.c_int32)),
    ]


class Timings(C.Structure):
    _fields_ = [
        ("pack_ms", C.c_double), ("h2d_ms", C.c_double), ("filter_ms", C.c_double), ("scan_ms", C.c_double),
        ("locate_ms", C.c_double), ("triage_ms", C.c_double), ("d2h_ms", C.c_double), ("total_ms", C.c_double),
        ("bases", C.c_int64), ("packed_bytes", C.c_int64), ("window_bytes", C.c_int64),
        ("h2d_bytes", C.c_int64), ("d2h_bytes", C.c_int64),
        ("kernel_launches", C.c_int32), ("scan_is_jit", C.c_int32), ("steps", C.c_int32), ("candidates", C.c_int32),
    ]


class Stage(C.Structure):
    _fields_ = [
        ("coarse_start", C.c_int32), ("coarse_end", C.c_int32), ("acc_start", C.c_int32), ("acc_end", C.c_int32),
        ("edge_start", C.c_int32), ("edge_end", C.c_int32), ("acc_density", C.c_double),
    ]


TRACK_DTYPE = np.dtype([("start", "<i4"), ("end", "<i4"), ("density", "<f8")])
RESULT_DTYPE = np.dtype([("status", "<i4"), ("n_win", "<i4"), ("track", TRACK_DTYPE, 3), ("win_offset", "<i8")])
assert RESULT_DTYPE.itemsize == 64

# every symbol include/nanotel_b200.h declares
EXPORTS = [
    "ntl_version", "ntl_create", "ntl_destroy", "ntl_last_error", "ntl_scan_batch", "ntl_scan_batch_concat",
    "ntl_scan_batch_pool", "ntl_scan_path", "ntl_scan_path_note", "ntl_device_count", "ntl_get_shards", "ntl_get_geometry",
    "ntl_jit_precompile_to", "ntl_jit_get_source",
    "ntl_batch_pack", "ntl_batch_upload", "ntl_batch_run", "ntl_batch_enqueue", "ntl_batch_wait", "ntl_batch_download", "ntl_get_timings", "ntl_stream",
    "ntl_get_windows", "ntl_get_window_counts", "ntl_get_stages", "ntl_jit_compile_check", "ntl_pack_read", "ntl_host_read_gbs", "ntl_write_read_outputs", "ntl_write_fasta_gz", "ntl_assign_serials", "ntl_count_windows",
    "ntl_reader_open", "ntl_reader_next", "ntl_reader_error", "ntl_reader_close",
]

_lib = None


def load() -> C.CDLL:
    """dlopen the library (no CUDA call happens here)."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise ImportError(
            "libnanotel_b200.so is not built: run `python telomere-analyzer_b200/build.py` "
            "(or __graft_entry__.build()).  There is no CPU fallback.")
    L = C.CDLL(LIB_PATH)
    vp, i32, i64 = C.c_void_p, C.c_int32, C.c_int64
    L.ntl_version.restype = C.c_int
    L.ntl_create.argtypes = [C.POINTER(vp), C.POINTER(Params)]
    L.ntl_destroy.argtypes = [vp]
    L.ntl_destroy.restype = None
    L.ntl_last_error.argtypes = [vp]
    L.ntl_last_error.restype = C.c_char_p
    L.ntl_scan_batch.argtypes = [vp, vp, vp, i32, C.POINTER(vp)]
    L.ntl_scan_batch_concat.argtypes = [vp, vp, vp, i32, C.POINTER(vp)]
    L.ntl_scan_batch_pool.argtypes = [vp, vp, vp, vp, i32, i32, C.POINTER(vp)]
    L.ntl_scan_path.argtypes = [vp]
    L.ntl_scan_path_note.argtypes = [vp]
    L.ntl_scan_path_note.restype = C.c_char_p
    L.ntl_device_count.argtypes = [vp]
    L.ntl_get_shards.argtypes = [vp, vp, vp, i32]
    L.ntl_get_geometry.argtypes = [vp, C.POINTER(i32), C.POINTER(i32), C.POINTER(i32), C.POINTER(i32)]
    L.ntl_jit_precompile_to.argtypes = [C.POINTER(Params), C.c_char_p, C.c_char_p, C.c_char_p, C.c_int]
    L.ntl_jit_precompile_to.restype = C.c_long
    L.ntl_jit_get_source.argtypes = [C.POINTER(Params), C.c_char_p, C.c_long]
    L.ntl_jit_get_source.restype = C.c_long
    L.ntl_batch_pack.argtypes = [vp, vp, vp, i32]
    L.ntl_batch_upload.argtypes = [vp]
    L.ntl_batch_run.argtypes = [vp]
    L.ntl_batch_enqueue.argtypes = [vp]
    L.ntl_batch_wait.argtypes = [vp]
    L.ntl_batch_download.argtypes = [vp, C.POINTER(vp)]
    L.ntl_get_timings.argtypes = [vp, C.POINTER(Timings)]
    L.ntl_stream.argtypes = [vp]
    L.ntl_stream.restype = vp
    L.ntl_get_windows.argtypes = [vp, i32, i32, i32, vp, vp, vp, vp]
    L.ntl_get_window_counts.argtypes = [vp, i32, vp, i64]
    L.ntl_get_window_counts.restype = i64
    L.ntl_get_stages.argtypes = [vp, i32, i32, C.POINTER(Stage)]
    L.ntl_jit_compile_check.argtypes = [C.POINTER(Params), C.c_char_p, C.c_char_p, C.c_int, C.c_char_p]
    L.ntl_jit_compile_check.restype = C.c_long
    L.ntl_pack_read.argtypes = [C.c_char_p, i64, i32, vp, i64, C.POINTER(i32)]
    L.ntl_pack_read.restype = C.c_long
    L.ntl_write_read_outputs.argtypes = [vp, C.c_char_p, vp, vp, vp, vp, vp, vp, i32, i32, C.c_double, i32, i32]
    L.ntl_write_read_outputs.restype = C.c_int
    L.ntl_write_fasta_gz.argtypes = [C.c_char_p, C.c_char_p, C.c_char_p, i64, i32]
    L.ntl_write_fasta_gz.restype = C.c_int
    L.ntl_host_read_gbs.argtypes = [vp, i64, i32, i32]
    L.ntl_host_read_gbs.restype = C.c_double
    L.ntl_assign_serials.argtypes = [vp, i32, i32, vp, vp, C.POINTER(i32)]
    L.ntl_count_windows.argtypes = [i64, i32]
    L.ntl_count_windows.restype = i32
    L.ntl_reader_open.argtypes = [C.POINTER(vp), C.POINTER(C.c_char_p), i32, C.c_char_p]
    L.ntl_reader_next.argtypes = [vp, i32, C.POINTER(vp), C.POINTER(vp), C.POINTER(vp), C.POINTER(vp)]
    L.ntl_reader_next.restype = i32
    L.ntl_reader_error.argtypes = [vp]
    L.ntl_reader_error.restype = C.c_char_p
    L.ntl_reader_close.argtypes = [vp]
    L.ntl_reader_close.restype = None
    _lib = L
    return L


def make_params(patterns, tvr_patterns=None, min_density=0.6, subseq_length=100, rc=False, use_filter=False,
                right_edge=False, device=0, options=0, host_threads=0, devices=None) -> Params:
    """patterns / tvr_patterns: a whitespace separated string (as on NanoTel.R's command line) or a sequence.
    devices: CUDA ordinals to shard every batch over (None: the single `device`)."""
    def toks(x):
        if x is None:
            return []
        if isinstance(x, (str, bytes)):
            x = x.split()
        return [t.encode() if isinstance(t, str) else bytes(t) for t in x]

    pats, tvr = toks(patterns), toks(tvr_patterns)
    P = Params()
    P.n_patterns = len(pats)
    arr_p = (C.c_char_p * max(len(pats), 1))(*pats)
    arr_t = (C.c_char_p * max(len(tvr), 1))(*tvr)
    P.patterns = C.cast(arr_p, C.POINTER(C.c_char_p))
    P.n_tvr = len(tvr)
    P.tvr_patterns = C.cast(arr_t, C.POINTER(C.c_char_p))
    P.min_density = float(min_density)
    P.subseq_length = int(subseq_length)
    P.rc = int(bool(rc))
    P.use_filter = int(bool(use_filter))
    P.right_edge = int(bool(right_edge))
    P.device = int(device)
    P.options = int(options)
    P.host_threads = int(host_threads)
    ids = [int(d) for d in devices] if devices else []
    arr_d = (C.c_int32 * max(len(ids), 1))(*ids)
    P.n_devices = len(ids)
    P.device_ids = C.cast(arr_d, C.POINTER(C.c_int32))
    P._keepalive = (pats, tvr, arr_p, arr_t, arr_d)
    return P
