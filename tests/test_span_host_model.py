"""CPU: the span scan kernel's arithmetic, checked without a GPU.

The source text the library generates for NVRTC (pattern tables + #include "ntl_scan.cuh") is compiled here with g++
as a host model (tests/host_model/span_model.cpp) and run over reads packed and laid out in spans exactly as the
batch packer does; every block count of every track must equal the oracle's window counts.  Covers the main pass
(interior spans: no validity masks), the tail pass, first-span / out-of-bounds behaviour at both ends, reads shorter
than one span, reads with N (four planes), --rc, three tracks, and block sizes that differ from the window size.
"""
import ctypes as C
import os
import subprocess

import numpy as np
import pytest

from oracle import oracle as O

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CSRC = os.path.join(ROOT, "telomere-analyzer_b200", "csrc")
FIRST, TAIL, SKIP = 1, 2, 4


def _build_model(tmp, patterns, tvr, S):
    from nanotel_b200 import _lib
    L = _lib.load()
    P = _lib.make_params(patterns, tvr, subseq_length=S)
    buf = C.create_string_buffer(1 << 20)
    n = L.ntl_jit_get_source(C.byref(P), buf, 1 << 20)
    assert n > 0, n
    src = os.path.join(tmp, "gen.cu.h")
    open(src, "wb").write(buf.value)
    so = os.path.join(tmp, "model.so")
    subprocess.run(["g++", "-O1", "-std=c++17", "-shared", "-fPIC", "-w", "-I", CSRC,
                    '-DNTL_GENERATED_SOURCE="%s"' % src, os.path.join(ROOT, "tests", "host_model", "span_model.cpp"),
                    "-o", so], check=True)
    M = C.CDLL(so)
    W, SG, T = C.c_int(), C.c_int(), C.c_int()
    M.span_model_geometry(C.byref(W), C.byref(SG), C.byref(T))
    return M, W.value, SG.value, T.value


def _layout(seqs, rc, W, BPS):
    """The batch packer's layout (ntl_api.cpp dev_pack): span-aligned reads, flags, 2-bit / 4-bit arenas."""
    from nanotel_b200 import _lib
    L = _lib.load()
    align = 8 // np.gcd(BPS, 8)
    spans_of = lambda n: (n + 32 * W - 1) // (32 * W)
    rup = lambda x, m: (x + m - 1) // m * m
    n = len(seqs)
    lens = np.array([len(s) for s in seqs], np.int32)
    packed, four = [], []
    for s in seqs:
        cap = ((len(s) + 31) // 32) * 4 + 8
        w = np.zeros(cap, np.uint32)
        fb = C.c_int32()
        k = L.ntl_pack_read(s, len(s), int(rc), w.ctypes.data, cap, C.byref(fb))
        assert k > 0
        packed.append(w[:k].copy()); four.append(bool(fb.value))
    woff = np.zeros(n, np.int64)
    cnt_off = np.zeros(n, np.int64)
    fmt = np.array(four, np.uint8)
    arenas, flags, nspans = [], [], []
    base_spans = 0
    for npl, sel in ((2, False), (4, True)):
        sp = 0
        first = {}
        for i in range(n):
            if four[i] == sel:
                first[i] = sp
                sp += rup(spans_of(int(lens[i])), align)
        tot = rup(sp, 32)
        lead = 8
        a = np.zeros(lead + (tot + 1) * W * npl, np.uint32)
        a[lead + tot * W * npl:] = 0xdeadbeef                  # the word after the last span: never trusted
        fl = np.full(max(tot, 1), SKIP, np.uint8)
        for i, s0 in first.items():
            a[lead + s0 * W * npl: lead + s0 * W * npl + len(packed[i])] = packed[i]
            nsp = spans_of(int(lens[i]))
            ltail = int(lens[i]) - (nsp - 1) * 32 * W
            tail_from = nsp - 2 if (nsp >= 2 and ltail < 18) else nsp - 1
            for s in range(nsp):
                fl[s0 + s] = (FIRST if s == 0 else 0) | (TAIL if s >= tail_from else 0)
            woff[i] = s0 * W
            cnt_off[i] = (base_spans + s0) * BPS
        arenas.append((a, lead)); flags.append(fl); nspans.append(tot)
        if not sel:
            base_spans = tot
    return lens, woff, cnt_off, fmt, arenas, flags, nspans


def _run_and_compare(tmp, seqs, patterns, tvr, S, rc):
    M, W, SG, T = _build_model(tmp, patterns, tvr, S)
    BPS = 32 * W // SG
    Q = S // SG
    assert S % SG == 0 and (32 * W) % SG == 0
    lens, woff, cnt_off, fmt, arenas, flags, nspans = _layout(seqs, rc, W, BPS)
    total = (nspans[0] + nspans[1]) * BPS + 8
    cnt = [np.full(total, 0xffff, np.uint16) for _ in range(3)]
    (a2, l2), (a4, l4) = arenas
    M.span_model_run.argtypes = [C.c_void_p, C.c_int64, C.c_void_p, C.c_void_p, C.c_int64, C.c_void_p, C.c_int,
                                 C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]
    # class bytes: one per span and track when a span holds at most 8 blocks; bit j <=> block j holds >= blk_thr bases
    blk_thr = max(1, (int(np.ceil(0.6 * S)) + Q - 1) // Q)
    cls = [np.full(nspans[0] + nspans[1] + 1, 0xff, np.uint8) for _ in range(3)]
    M.span_model_set_cls.argtypes = [C.c_uint32, C.c_void_p, C.c_void_p, C.c_void_p]
    M.span_model_set_cls(blk_thr, cls[0].ctypes.data, cls[1].ctypes.data, cls[2].ctypes.data)
    M.span_model_run(a2.ctypes.data + 4 * l2, nspans[0], flags[0].ctypes.data, a4.ctypes.data + 4 * l4, nspans[1],
                     flags[1].ctypes.data, len(seqs), lens.ctypes.data, woff.ctypes.data, fmt.ctypes.data,
                     cnt[0].ctypes.data, cnt[1].ctypes.data, cnt[2].ctypes.data)
    P = O.make_params(patterns, tvr, 0.6, S, False)
    recs, passed, win_off, wc = O.scan_batch(P, seqs, do_rc=rc, n_threads=4)
    bad = []
    for i, s in enumerate(seqs):
        nw = int(recs[i]["n_win"])
        nb = (len(s) + SG - 1) // SG
        for t in range(T):
            blk = cnt[t][cnt_off[i]: cnt_off[i] + nb].astype(np.int64)
            got = [int(blk[k * Q: (nb if k == nw - 1 else (k + 1) * Q)].sum()) for k in range(nw)]
            exp = wc[int(win_off[i]) + t * nw: int(win_off[i]) + (t + 1) * nw].tolist()
            if got != exp:
                k = [j for j in range(nw) if got[j] != exp[j]][0]
                bad.append((i, len(s), t, k, got[k], exp[k]))
    assert not bad, bad[:8]
    if BPS <= 8:
        for i, s in enumerate(seqs):
            nb = (len(s) + SG - 1) // SG
            for t in range(T):
                blk = cnt[t][cnt_off[i]: cnt_off[i] + nb].astype(np.int64)
                by = cls[t][cnt_off[i] // BPS: cnt_off[i] // BPS + (nb + BPS - 1) // BPS]
                got = np.array([(int(by[j // BPS]) >> (j % BPS)) & 1 for j in range(nb)], bool)
                assert (got == (blk >= blk_thr)).all(), (i, len(s), t)
    return W, SG


CASES = [
    ("TTAGGG", None, 100, False),
    ("YYAGGG", None, 100, True),
    ("YYAGGG", "TTGGG CCAGGG TCAGGG", 100, True),
    ("TTAGGG", None, 200, False),
    ("TTAGGG", None, 500, False),
    ("TTAGGG TTGGG", None, 100, False),
    ("RRTCCC", "CCCAA", 64, True),
    ("TTAGGG", None, 20, False),
    ("TTAGGGTTAGGGTTAGGG", "NNAGGG", 150, False),
    ("GGG", None, 48, False),
]


@pytest.mark.parametrize("patterns,tvr,S,rc", CASES, ids=lambda v: str(v).replace(" ", "+"))
def test_span_model_equals_oracle(tmp_path, example_reads, patterns, tvr, S, rc):
    from nanotel_b200.synth import as_list, synth_reads
    rng = np.random.default_rng(abs(hash((patterns, S))) % (1 << 31))
    seqs = [s for _, s in example_reads]
    buf, off, _ = synth_reads(60, 77, telomeric_frac=0.5, median_len=2500, max_len=9000, min_len=300, n_frac=0.3)
    seqs += as_list(buf, off)
    # lengths around every span / block / word boundary, ends made of repeats so that both out-of-bounds rules fire
    unit = (patterns.split()[0] * 40).replace("Y", "T").replace("R", "A").replace("N", "C").encode()
    for Ln in list(range(1, 40)) + [63, 64, 65, 99, 100, 101, 150, 151, 199, 200, 201, 767, 768, 769, 799, 800, 801,
                                    815, 816, 817, 818, 832, 1599, 1600, 1601, 1617, 1618, 2400, 3200, 3201]:
        body = bytes(rng.choice(np.frombuffer(b"ACGT", np.uint8), Ln))
        k = int(rng.integers(0, 6))
        seqs.append((unit[k:k + Ln // 2] + body)[:Ln])
        seqs.append((body + unit[k:k + 30])[-Ln:])
        seqs.append((unit * 30)[k:k + Ln])
    W, SG = _run_and_compare(str(tmp_path), seqs, patterns, tvr, S, rc)
    assert 8 <= W <= 32 and S % SG == 0
