"""GPU: inputs at the edges of the implementation's internal limits, compared with the oracle bit for bit.

* very long reads: more windows than the triage kernel walks (8192) and than the locate kernel keeps as class bits in
  shared memory (16384) -> the global-memory fallback of the run/score machine;
* an empty batch, a batch of one 1-base read, call-order errors."""
import numpy as np
import pytest

from helpers import compare_batch, oracle_batch

pytestmark = pytest.mark.gpu
ACGT = np.frombuffer(b"ACGT", np.uint8)


def _long_read(rng, L, telo_at_end, telo_len=12000, unit=b"TTAGGG"):
    s = bytearray(rng.choice(ACGT, L).tobytes())
    rep = bytearray((unit * (telo_len // len(unit) + 1))[:telo_len])
    for _ in range(telo_len // 40):
        rep[int(rng.integers(0, telo_len))] = ACGT[int(rng.integers(0, 4))]
    if telo_at_end:
        s[L - telo_len - 30:L - 30] = rep
    else:
        s[25:25 + telo_len] = rep
    # a few isolated telomeric windows in the middle (runs that never qualify)
    for pos in (L // 3, L // 2, L // 2 + 350):
        s[pos:pos + 150] = (unit * 25)[:150]
    return bytes(s)


@pytest.mark.parametrize("S,L", [(100, 900_000), (100, 1_700_000), (20, 400_000), (500, 3_000_000)])
def test_very_long_reads(S, L):
    from nanotel_b200 import Scanner
    rng = np.random.default_rng(L + S)
    seqs = [_long_read(rng, L, True), _long_read(rng, L - 777, False), bytes(rng.choice(ACGT, L // 2)),
            _long_read(rng, 5000, True, 2000)]
    for right in (False, True):
        P, recs, passed, win_off, wc = oracle_batch(seqs, "TTAGGG", "TTGGG", 0.6, S, right, False, False, n_threads=4)
        with Scanner("TTAGGG", "TTGGG", 0.6, S, right_edge=right, debug_stages=True) as sc:
            res = sc.scan(seqs)
            compare_batch(sc, res, seqs, recs, passed, win_off, wc, check_stages=True, label="S=%d L=%d right=%s" % (S, L, right))
            assert res[0]["status"] & 1 and res[1]["status"] & 1 and not (res[2]["status"] & 1)


def test_empty_batch_and_tiny_reads():
    from nanotel_b200 import READ_FILTERED, READ_KEEP, READ_REF_ERROR, NanoTelError, Scanner
    with Scanner("TTAGGG", use_filter=True) as sc:
        res = sc.scan([b"", b"ACGT" * 300, b""])                    # width < 1000: filter_reads drops them (:2124)
        assert all(r["status"] & READ_FILTERED for r in res) and not any(r["status"] & READ_REF_ERROR for r in res)
    with Scanner("TTAGGG") as sc:
        res = sc.scan([])
        assert len(res) == 0
        res = sc.scan([b"A"])
        assert len(res) == 1 and res[0]["n_win"] == 0 and not (res[0]["status"] & 1)
        # a zero-length read: NanoTel.R stops on it (seq(1, 0, by = S), :216) unless filter_reads dropped it first
        res = sc.scan([b"ACGT", b"", b"TTAGGG" * 50])
        assert res[1]["status"] & READ_REF_ERROR and not (res[1]["status"] & READ_KEEP)
        assert res[2]["status"] & READ_KEEP and not (res[0]["status"] & (READ_KEEP | READ_REF_ERROR))
        with pytest.raises(NanoTelError) as e:
            sc.scan([b"ACGTJ"])
        assert e.value.code == -3                                   # not a DNA letter
        res = sc.scan([b"TTAGGG" * 50])                             # the context is still usable
        assert res[0]["status"] & 1


def test_call_order_is_checked():
    from nanotel_b200 import NanoTelError, Scanner
    with Scanner("TTAGGG") as sc:
        with pytest.raises(NanoTelError) as e:
            sc.run()
        assert e.value.code == -7
        sc.pack([b"TTAGGG" * 40])
        with pytest.raises(NanoTelError):
            sc.download()
        sc.upload(); sc.run()
        res = sc.download()
        assert res[0]["status"] & 1


def test_only_kept_reads_send_their_window_tables_to_the_host():
    """ntl_batch_download moves the 64-byte records and the window prefixes of the kept reads; the table of any other
    read is fetched from the device when ntl_get_windows asks for it.  Both routes must give the same table as a
    context that holds only that read.  `out=` receives the records without an allocation."""
    from nanotel_b200 import RESULT_DTYPE, Scanner
    rng = np.random.default_rng(77)
    seqs = []
    for i in range(60):
        L = int(rng.integers(1500, 9000))
        seqs.append(_long_read(rng, L, bool(i % 3 == 0), 1200) if i % 2 == 0 else bytes(rng.choice(ACGT, L)))
    out = np.empty(len(seqs) + 5, RESULT_DTYPE)
    with Scanner("TTAGGG", "TTGGG") as sc:
        res = sc.scan(seqs, out=out)
        assert np.shares_memory(res, out) and len(res) == len(seqs)
        assert res.tobytes() == sc.scan(seqs).tobytes()
        view = sc.scan(seqs, out="view")                       # the C ABI's own buffer: no copy, read-only
        assert view.tobytes() == res.tobytes() and not view.flags.writeable and not view.flags.owndata
        with pytest.raises(ValueError):
            sc.scan(seqs, out="copy")
        tm = sc.timings()
        kept = np.flatnonzero(res["status"] & 1)
        assert 0 < len(kept) < len(seqs)
        blk = sc.geometry()["block"]                               # count blocks per read, rows padded to 8 entries
        rows = sum(((-(-len(seqs[i]) // blk) + 7) & ~7) for i in kept)
        assert tm["d2h_bytes"] == 64 * len(seqs) + rows * 2 * sc.n_tracks
        tables = {(i, t): sc.windows(i, t) for i in range(len(seqs)) for t in range(sc.n_tracks)}
        with pytest.raises(ValueError):
            sc.scan(seqs, out=np.empty(3, RESULT_DTYPE))
    with Scanner("TTAGGG", "TTGGG") as one:
        for i in list(kept[:4]) + [j for j in range(len(seqs)) if j not in set(kept)][:4]:
            one.scan([seqs[i]])
            for t in range(one.n_tracks):
                for a, b in zip(tables[(i, t)], one.windows(0, t)):
                    assert np.array_equal(a, b), (i, t)


@pytest.mark.parametrize("patterns", ["TTAGGG", "TTAGGG CCCTAA", "TTAGGG GGGTTA", "ACACAC", "YYAGGG", "TTAGGG TTAGG CCCTAA",
                                      "AAAAAA", "TTAGGG TTAGGC"])
def test_exact_coverage_by_subtraction_only_where_hits_cannot_overlap(patterns):
    """The NVRTC kernel covers the exact hits of an 'unbordered' group of equal-length patterns with (H << m) - H and
    every other group by dilation; both must give the oracle's coverage on reads made of abutting, overlapping and
    phase-shifted repeats of the patterns themselves (hits at word and lane-block boundaries included)."""
    from nanotel_b200 import Scanner
    rng = np.random.default_rng(len(patterns) * 7919)
    toks = [t.replace("Y", "C").encode() for t in patterns.split()]
    seqs = []
    for i in range(24):
        parts, total = [], 0
        while total < 300 + 611 * i:
            u = toks[int(rng.integers(0, len(toks)))]
            k = int(rng.integers(1, 40))
            piece = (u * k)[int(rng.integers(0, len(u))):]
            if rng.random() < 0.3:
                piece = piece + bytes(rng.choice(ACGT, int(rng.integers(1, 9))))
            parts.append(piece); total += len(piece)
        seqs.append(b"".join(parts))
    P, recs, passed, win_off, wc = oracle_batch(seqs, patterns, None, 0.6, 100, False, False, False)
    for jit in (True, False):
        with Scanner(patterns, None, 0.6, 100, jit=jit, debug_stages=True) as sc:
            res = sc.scan(seqs)
            compare_batch(sc, res, seqs, recs, passed, win_off, wc, check_stages=True, label="%s jit=%s" % (patterns, jit))
