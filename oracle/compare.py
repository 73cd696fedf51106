"""Whole-batch, vectorised comparison of the CUDA path with the oracle -- TEST INFRASTRUCTURE, NOT PRODUCT.

Used by tests/ (full-size parity: every record and every window count of 100 000+ reads) and by bench.py's parity
leg, which runs it on the timed workload outside the timed region.  The product package never imports this module.

    v = full_parity(scanner, results, seqs_or_(buf, offsets), patterns, tvr, min_density, S, right_edge, rc, use_filter)
    v == {"reads": n, "reads_compared": ..., "windows": ..., "mismatches": 0, ...}
"""
from __future__ import annotations

import time
from typing import Optional, Sequence

import numpy as np

from . import oracle as O

READ_KEEP, READ_FILTERED, READ_REF_ERROR = 1, 2, 4


def _as_seqs(reads) -> Sequence[bytes]:
    if isinstance(reads, tuple):
        buf, off = reads
        return [buf[int(off[i]):int(off[i + 1])].tobytes() for i in range(len(off) - 1)]
    return reads


def compare_records(res: np.ndarray, recs: np.ndarray, passed: np.ndarray, T: int):
    """Bit-exact comparison of the per-read records.  Returns (bad_read_mask, comparable_mask, detail dict)."""
    n = len(res)
    assert len(recs) == n and len(passed) == n
    st = res["status"].astype(np.int64)
    filt_g = (st & READ_FILTERED) != 0
    filt_o = passed == 0
    bad = filt_g != filt_o
    detail = {"filter_verdict": int(bad.sum())}
    live = ~filt_g & ~filt_o
    err_g = (st & READ_REF_ERROR) != 0
    err_o = (recs["flags"] & O.FLAG_REF_ERROR) != 0
    b = live & (err_g != err_o)
    detail["ref_error"] = int(b.sum()); bad |= b
    cmpm = live & ~err_g & ~err_o
    b = cmpm & (res["n_win"] != recs["n_win"])
    detail["n_win"] = int(b.sum()); bad |= b
    b = cmpm & (((st & READ_KEEP) != 0) != (recs["keep"] != 0))
    detail["keep"] = int(b.sum()); bad |= b
    for t in range(T):
        g, e = res["track"][:, t], recs["t"][:, t]
        b = cmpm & ((g["start"] != e["start"]) | (g["end"] != e["end"]))
        detail["track%d_interval" % t] = int(b.sum()); bad |= b
        gd = np.ascontiguousarray(g["density"]).view(np.uint64)
        ed = np.ascontiguousarray(e["density"]).view(np.uint64)
        b = cmpm & (gd != ed)
        detail["track%d_density_bits" % t] = int(b.sum()); bad |= b
    return bad, cmpm, detail


def compare_windows(scanner, res: np.ndarray, recs: np.ndarray, cmpm: np.ndarray, win_off: np.ndarray, wc: np.ndarray,
                    T: int):
    """Every window count of every comparable read, all tracks.  Returns (windows compared, mismatching windows,
    reads with a mismatching window)."""
    nw = np.where(cmpm, recs["n_win"], 0).astype(np.int64)
    g_nw = res["n_win"].astype(np.int64)
    g_off = np.zeros(len(res) + 1, np.int64)
    np.cumsum(g_nw, out=g_off[1:])
    ok_reads = cmpm & (g_nw == recs["n_win"])
    nw = np.where(ok_reads, nw, 0)
    total = int(nw.sum())
    if total == 0:
        return 0, 0, 0
    rid = np.repeat(np.arange(len(res)), nw)
    within = np.arange(total) - np.repeat(np.cumsum(nw) - nw, nw)
    n_bad_w, bad_reads = 0, np.zeros(len(res), bool)
    for t in range(T):
        g_all = scanner.window_counts(t, int(g_off[-1]))
        g = g_all[g_off[rid] + within].astype(np.int32)
        e = wc[win_off[rid] + t * nw[rid] + within]
        d = g != e
        n_bad_w += int(d.sum())
        if d.any():
            bad_reads[rid[d]] = True
    return total * T, n_bad_w, int(bad_reads.sum())


def full_parity(scanner, res: np.ndarray, reads, patterns, tvr=None, min_density: float = 0.6, S: int = 100,
                right_edge: bool = False, rc: bool = False, use_filter: bool = False, n_threads: int = 0,
                check_windows: bool = True) -> dict:
    """Run the oracle over the same reads and compare everything.  `res` = the records the CUDA path returned for the
    scanner's LAST batch (the window counts are fetched from that batch)."""
    import os
    seqs = _as_seqs(reads)
    T = 3 if tvr else 2
    P = O.make_params(patterns, tvr, min_density, S, right_edge)
    t0 = time.perf_counter()
    recs, passed, win_off, wc = O.scan_batch(P, seqs, do_rc=rc, use_filter=use_filter,
                                             n_threads=n_threads or (os.cpu_count() or 1), want_windows=check_windows)
    t_oracle = time.perf_counter() - t0
    bad, cmpm, detail = compare_records(res, recs, passed, T)
    out = {"reads": int(len(res)), "reads_compared": int(cmpm.sum()), "reads_filtered": int((passed == 0).sum()),
           "reads_kept": int(((res["status"] & READ_KEEP) != 0).sum()),
           "record_mismatches": int(bad.sum()), "windows": 0, "window_mismatches": 0}
    if check_windows:
        nwin, nbad, nbad_reads = compare_windows(scanner, res, recs, cmpm, win_off, wc, T)
        out["windows"] = nwin
        out["window_mismatches"] = nbad
        detail["reads_with_window_mismatch"] = nbad_reads
    out["mismatches"] = out["record_mismatches"] + out["window_mismatches"]
    out["oracle_s"] = round(t_oracle, 2)
    if out["mismatches"]:
        out["detail"] = {k: v for k, v in detail.items() if v}
        out["first_bad_reads"] = [int(i) for i in np.nonzero(bad)[0][:5]]
    return out
