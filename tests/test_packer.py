"""CPU: the host packer (AVX2 + scalar tails, --rc folded in, 4-bit fallback) against a plain numpy restatement of
the layout in csrc/ntl_dev.h: position p (1-based, after rc) is bit (p - 1) of the read's stream; every 32 positions
are one record {lo, hi} with code (ASCII >> 1) & 3, or {A, C, G, T} for reads with IUPAC letters."""
import ctypes as C

import numpy as np
import pytest

NIB = {c: v for c, v in zip(b"ACGTMRWSYKVHDBN", [1, 2, 4, 8, 3, 5, 9, 6, 10, 12, 7, 11, 13, 14, 15])}
NIB.update({ord("-"): 0, ord("+"): 0, ord("."): 0})


def _pack(seq: bytes, rc: bool, misalign: int = 0):
    """misalign = 0: a 32-byte aligned destination, as the batch buffers are (the AVX2 path then uses non-temporal
    stores); otherwise a destination `misalign` words off that alignment (plain stores)."""
    from nanotel_b200 import _lib
    L = _lib.load()
    cap = ((len(seq) >> 5) + 8) * 4 * 4
    raw = np.zeros(cap + 16, np.uint32)
    skip = ((-raw.ctypes.data) % 32) // 4 + misalign
    w = raw[skip:skip + cap]
    assert (w.ctypes.data % 32 == 0) == (misalign == 0)
    fb = C.c_int32()
    n = L.ntl_pack_read(seq, len(seq), int(rc), w.ctypes.data, cap, C.byref(fb))
    assert n > 0, n
    return w[:n], bool(fb.value)


def _expected(seq: bytes, rc: bool):
    from oracle import oracle as O
    s = O.revcomp(seq) if rc else seq.upper()
    n_words = (len(s) + 31) >> 5
    acgt = all(c in b"ACGT" for c in s)
    planes = 2 if acgt else 4
    bits = np.zeros((planes, n_words * 32), np.uint8)
    for p, c in enumerate(s):
        if acgt:
            code = (c >> 1) & 3
            bits[0, p] = code & 1
            bits[1, p] = code >> 1
        else:
            nb = NIB[c]
            for k in range(4):
                bits[k, p] = (nb >> k) & 1
    words = np.zeros(n_words * planes, np.uint32)
    for k in range(planes):
        words[k::planes] = np.packbits(bits[k].reshape(-1, 32), axis=1, bitorder="little").view(np.uint32).ravel()
    return words, not acgt


@pytest.mark.parametrize("rc", [False, True])
def test_packer_matches_layout(rc):
    rng = np.random.default_rng(12)
    lens = list(range(1, 70)) + [95, 96, 97, 127, 128, 129, 255, 256, 257, 383, 384, 385, 1000, 4095, 4096, 4097, 10007]
    for L in lens:
        seq = bytes(rng.choice(np.frombuffer(b"ACGTacgt", np.uint8), L))
        got, fb = _pack(seq, rc, misalign=L % 3)
        exp, efb = _expected(seq, rc)
        assert not fb and not efb
        assert np.array_equal(got, exp), (L, rc)
    # every 8-byte alignment of the output inside a 32-byte block (the vector loop starts at the first 32-byte boundary
    # and leaves 0-3 head words and 0-3 tail words + a partial word to the block code), lengths around every hand-over
    for mis in (0, 2, 4, 6):
        for L in list(range(120, 140)) + list(range(155, 165)) + list(range(250, 262)) + list(range(284, 292)) + \
                list(range(380, 390)) + list(range(509, 518)) + [1000, 4095, 4096, 4097, 4099, 4127, 4128, 4129]:
            seq = bytes(rng.choice(np.frombuffer(b"ACGTacgt", np.uint8), L))
            got, fb = _pack(seq, rc, misalign=mis)
            exp, efb = _expected(seq, rc)
            assert not fb and not efb and np.array_equal(got, exp), (L, rc, mis)
    for L in (1, 5, 33, 128, 129, 700):
        seq = bytearray(rng.choice(np.frombuffer(b"ACGT", np.uint8), L))
        seq[int(rng.integers(0, L))] = ord("N")
        if L > 40:
            seq[37] = ord("r")
            seq[L - 1] = ord("-")
        got, fb = _pack(bytes(seq), rc)
        exp, efb = _expected(bytes(seq), rc)
        assert fb and efb and np.array_equal(got, exp), (L, rc)


@pytest.mark.parametrize("rc", [False, True])
def test_iupac_reads_take_the_vector_path_and_fall_back_on_gaps(rc):
    """4-bit reads: 32-letter blocks of IUPAC letters are packed by the AVX2 loop, the first block holding a gap
    character (or the tail) hands over to the scalar loop with the carried bits; every hand-over point is tried."""
    rng = np.random.default_rng(99)
    iupac = np.frombuffer(b"ACGTMRWSYKVHDBNacgtmrwsykvhdbn", np.uint8)
    for L in list(range(1, 100)) + [127, 128, 129, 1000, 4097]:
        seq = bytearray(rng.choice(iupac, L))
        seq[int(rng.integers(0, L))] = ord("N")                         # at least one non-ACGT letter
        got, fb = _pack(bytes(seq), rc, misalign=L % 2)
        exp, efb = _expected(bytes(seq), rc)
        assert fb and efb and np.array_equal(got, exp), (L, rc)
    for gap_at in (0, 31, 32, 63, 64, 100, 299):
        seq = bytearray(rng.choice(iupac, 300))
        seq[gap_at] = ord("-+."[gap_at % 3])
        got, fb = _pack(bytes(seq), rc)
        exp, efb = _expected(bytes(seq), rc)
        assert fb and efb and np.array_equal(got, exp), (gap_at, rc)
    for bad_at in (5, 40, 170):                                         # a letter outside the alphabet, any block
        seq = bytearray(rng.choice(iupac, 200))
        seq[3] = ord("N")
        seq[bad_at] = ord("J")
        from nanotel_b200 import _lib
        L_ = _lib.load()
        w = np.zeros(4096, np.uint32)
        fbv = C.c_int32()
        assert L_.ntl_pack_read(bytes(seq), len(seq), int(rc), w.ctypes.data, 4096, C.byref(fbv)) < 0


def test_packer_rejects_non_dna():
    from nanotel_b200 import _lib
    L = _lib.load()
    w = np.zeros(64, np.uint32)
    fb = C.c_int32()
    assert L.ntl_pack_read(b"ACGTXACGT", 9, 0, w.ctypes.data, 64, C.byref(fb)) == -3
    assert L.ntl_pack_read(b"ACGT", 0, 0, w.ctypes.data, 64, C.byref(fb)) == -1
