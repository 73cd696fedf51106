"""CPU: the whole-batch comparison used by the full-size GPU parity tests and bench.py's parity leg (oracle/compare.py)
reports zero mismatches on a faithful result set and finds every kind of planted difference."""
import numpy as np

from oracle import oracle as O
from oracle.compare import full_parity


class _FakeScanner:
    def __init__(self, counts):
        self.counts = counts

    def window_counts(self, track, total=None):
        return self.counts[track]


def _faithful(seqs, patterns, tvr, S, right_edge, rc, use_filter):
    from nanotel_b200 import _lib
    P = O.make_params(patterns, tvr, 0.6, S, right_edge)
    recs, passed, win_off, wc = O.scan_batch(P, seqs, do_rc=rc, use_filter=use_filter, n_threads=4)
    T = 3 if tvr else 2
    res = np.zeros(len(seqs), _lib.RESULT_DTYPE)
    res["n_win"] = recs["n_win"]
    res["status"] = np.where(passed == 0, 2, np.where(recs["flags"] & 1, 4, recs["keep"]))
    for t in range(T):
        for k in ("start", "end", "density"):
            res["track"][k][:, t] = recs["t"][k][:, t]
    counts = []
    for t in range(T):
        parts = [wc[int(win_off[i]) + t * int(recs["n_win"][i]): int(win_off[i]) + (t + 1) * int(recs["n_win"][i])]
                 for i in range(len(seqs))]
        counts.append(np.concatenate(parts).astype(np.uint16) if parts else np.zeros(0, np.uint16))
    return res, counts


def test_full_parity_accepts_faithful_results_and_finds_planted_differences():
    from nanotel_b200.synth import as_list, synth_reads
    buf, off, meta = synth_reads(300, 11, telomeric_frac=0.4, median_len=4000, max_len=30000)
    seqs = as_list(buf, off)
    args = ("YYAGGG", "TTGGG", 0.6, 100, True, True, True)
    res, counts = _faithful(seqs, "YYAGGG", "TTGGG", 100, True, True, True)
    v = full_parity(_FakeScanner(counts), res, seqs, *args)
    assert v["mismatches"] == 0 and v["reads"] == 300 and v["windows"] > 0 and 0 < v["reads_compared"] < 300, v
    v2 = full_parity(_FakeScanner(counts), res, (buf, off), *args)        # (buffer, offsets) form
    assert v2["mismatches"] == 0 and v2["windows"] == v["windows"]

    live = np.nonzero((res["status"] & 2) == 0)[0]
    kept = np.nonzero(res["status"] & 1)[0]
    assert len(kept) > 5
    r = res.copy(); r["track"]["end"][kept[0], 1] += 1
    assert full_parity(_FakeScanner(counts), r, seqs, *args)["record_mismatches"] == 1
    r = res.copy(); r["track"]["density"][kept[1], 2] = np.nextafter(r["track"]["density"][kept[1], 2], 2.0)
    v = full_parity(_FakeScanner(counts), r, seqs, *args)
    assert v["record_mismatches"] == 1 and v["detail"]["track2_density_bits"] == 1
    r = res.copy(); r["status"][kept[2]] &= ~1
    assert full_parity(_FakeScanner(counts), r, seqs, *args)["detail"]["keep"] == 1
    r = res.copy(); r["status"][live[0]] = 2
    assert full_parity(_FakeScanner(counts), r, seqs, *args)["detail"]["filter_verdict"] == 1
    c = [x.copy() for x in counts]; c[1][len(c[1]) // 2] ^= 1
    v = full_parity(_FakeScanner(c), res, seqs, *args)
    assert v["window_mismatches"] == 1 and v["record_mismatches"] == 0 and v["mismatches"] == 1
