"""GPU, BASELINE-sized inputs: size-independent properties of the hot path on the cfg2 workload (100 000 synthetic
ONT-like reads would take the oracle minutes; here 20 000 reads = 0.46 Gbases per check, and a 300-read sample is
compared with the oracle):

  * the NVRTC-specialised and the runtime-pattern scan kernel produce identical records and window prefixes;
  * --rc equals scanning the reverse-complemented reads without --rc;
  * results do not depend on how the reads are batched or ordered;
  * the per-window counts are consistent with the per-read densities (sum of window counts over the windows fully
    inside the called interval never exceeds the covered count the density implies);
  * a random sample agrees with the oracle bit for bit."""
import numpy as np
import pytest

from helpers import compare_batch, oracle_batch

pytestmark = pytest.mark.gpu
N = 20000


@pytest.fixture(scope="module")
def workload():
    from nanotel_b200.synth import synth_reads
    buf, off, meta = synth_reads(N, 20261018 + 2)
    return buf, off, meta


def _scan(buf, off, **kw):
    from nanotel_b200 import Scanner
    with Scanner("YYAGGG", kw.pop("tvr", None), 0.6, 100, **kw) as sc:
        res = sc.scan_concat(buf, off)
        cums = [np.concatenate([sc.windows(i, t)[2] for i in range(0, len(res), 997)]) for t in range(sc.n_tracks)]
    return res, cums


def test_jit_and_runtime_kernels_agree(workload):
    buf, off, _ = workload
    a, ca = _scan(buf, off, rc=True, jit=True)
    b, cb = _scan(buf, off, rc=True, jit=False)
    assert a.tobytes() == b.tobytes()
    for x, y in zip(ca, cb):
        assert np.array_equal(x, y)
    assert int((a["status"] & 1).sum()) > 500


def test_rc_flag_equals_scanning_reversed_reads(workload):
    from oracle import oracle as O
    buf, off, _ = workload
    n = 3000
    seqs = [buf[int(off[i]):int(off[i + 1])].tobytes() for i in range(n)]
    from nanotel_b200 import Scanner
    with Scanner("YYAGGG", "TTGGG CCAGGG TCAGGG", rc=True) as sc:
        a = sc.scan(seqs)
    with Scanner("YYAGGG", "TTGGG CCAGGG TCAGGG", rc=False) as sc:
        b = sc.scan([O.revcomp(s) for s in seqs])
    assert a.tobytes() == b.tobytes()


def test_batching_and_order_do_not_matter(workload):
    buf, off, _ = workload
    from nanotel_b200 import Scanner
    n = 6000
    seqs = [buf[int(off[i]):int(off[i + 1])].tobytes() for i in range(n)]
    perm = np.random.default_rng(0).permutation(n)
    with Scanner("YYAGGG", rc=True) as sc:
        whole = sc.scan(seqs)
        parts = np.concatenate([sc.scan(seqs[k:k + 1700]) for k in range(0, n, 1700)])
        shuffled = sc.scan([seqs[i] for i in perm])
    for name in ("status", "n_win", "track"):
        assert np.array_equal(whole[name], parts[name])
        assert np.array_equal(whole[name][perm], shuffled[name])


def test_windows_are_consistent_with_densities(workload):
    buf, off, meta = workload
    from nanotel_b200 import Scanner
    with Scanner("YYAGGG", rc=True) as sc:
        res = sc.scan_concat(buf, off)
        kept = np.nonzero(res["status"] & 1)[0][:400]
        for i in kept:
            L = int(meta["lengths"][i])
            for t in range(2):
                tr = res[i]["track"][t]
                s, e = int(tr["start"]), int(tr["end"])
                if s == -1:
                    continue
                st, en, cov, den = sc.windows(int(i), t)
                assert en[-1] == L and cov.max() <= (en - st + 1).max()
                covered = round(float(tr["density"]) * (e - s + 1))
                inside = (st >= s) & (en <= e)
                touching = (en >= s) & (st <= e)
                assert cov[inside].sum() <= covered <= cov[touching].sum()


def test_sample_agrees_with_the_oracle(workload):
    buf, off, _ = workload
    from nanotel_b200 import Scanner
    idx = np.random.default_rng(1).choice(N, 300, replace=False)
    seqs = [buf[int(off[i]):int(off[i + 1])].tobytes() for i in idx]
    P, recs, passed, win_off, wc = oracle_batch(seqs, "YYAGGG", "TTGGG CCAGGG TCAGGG", 0.6, 100, False, True, False, n_threads=8)
    with Scanner("YYAGGG", "TTGGG CCAGGG TCAGGG", rc=True, debug_stages=True) as sc:
        res = sc.scan(seqs)
        compare_batch(sc, res, seqs, recs, passed, win_off, wc, check_stages=True, label="cfg3 sample")
