"""GPU parity at BASELINE sizes: EVERY per-read record and EVERY window count of the full workloads against the oracle
(oracle/compare.py; the oracle scans 100 000 reads in a few seconds on the box's host cores).

  cfg2   100 000 reads / 2.3 Gbases, --patterns YYAGGG --rc                     (the timed workload of bench.py)
  cfg3   the same reads + --tvr_patterns "TTGGG CCAGGG TCAGGG" (three tracks)
  cfg4   >= 100 000 reads, --patterns TTAGGG --use_filter --check_right_edge, subseq_length 100 / 200 / 500
  cfg4N  the same with reads that carry N (4-bit path) under the filter
"""
import numpy as np
import pytest

from oracle.compare import full_parity

pytestmark = pytest.mark.gpu
SEED = 20261018


@pytest.fixture(scope="module")
def cfg2_reads():
    from nanotel_b200.synth import synth_reads
    return synth_reads(100000, SEED + 2)


def _check(v, min_kept):
    assert v["mismatches"] == 0, v
    assert v["reads_kept"] >= min_kept, v
    assert v["windows"] > 0


@pytest.mark.parametrize("tvr", [None, "TTGGG CCAGGG TCAGGG"], ids=["cfg2", "cfg3"])
def test_all_reads_of_cfg2_and_cfg3(cfg2_reads, tvr):
    from nanotel_b200 import Scanner
    buf, off, meta = cfg2_reads
    with Scanner("YYAGGG", tvr, 0.6, 100, rc=True) as sc:
        res = sc.scan_concat(buf, off)
        assert sc.timings()["scan_is_jit"] == 1
        v = full_parity(sc, res, (buf, off), "YYAGGG", tvr, 0.6, 100, False, True, False)
    assert v["reads"] == 100000 and v["reads_compared"] == 100000
    assert v["windows"] >= (3 if tvr else 2) * 23000000
    _check(v, 4000)


@pytest.fixture(scope="module")
def cfg4_reads():
    from nanotel_b200.synth import synth_reads
    # 30 % telomeric so that the right-edge filter keeps a five-digit number of reads
    return synth_reads(120000, SEED + 4, telomeric_frac=0.30, median_len=12000.0)


@pytest.mark.parametrize("S", [100, 200, 500])
def test_filter_right_edge_subseq_sweep(cfg4_reads, S):
    from nanotel_b200 import Scanner
    buf, off, meta = cfg4_reads
    with Scanner("TTAGGG", None, 0.6, S, use_filter=True, right_edge=True) as sc:
        res = sc.scan_concat(buf, off)
        v = full_parity(sc, res, (buf, off), "TTAGGG", None, 0.6, S, True, False, True)
    assert v["reads"] == 120000
    assert 10000 <= v["reads_compared"] <= 60000, v        # the reads the filter let through
    _check(v, 8000)


@pytest.mark.parametrize("S", [100, 500])
def test_filter_with_n_reads(S):
    """Reads with N (four nibble planes) under --use_filter --check_right_edge, --rc, ambiguity pattern + TVR."""
    from nanotel_b200 import READ_IUPAC, Scanner
    from nanotel_b200.synth import synth_reads
    buf, off, meta = synth_reads(30000, SEED + 44, telomeric_frac=0.4, median_len=8000.0, n_frac=0.5)
    with Scanner("YYAGGG", "TTGGG", 0.6, S, rc=True, use_filter=True, right_edge=True) as sc:
        res = sc.scan_concat(buf, off)
        v = full_parity(sc, res, (buf, off), "YYAGGG", "TTGGG", 0.6, S, True, True, True)
    assert int(((res["status"] & READ_IUPAC) != 0).sum()) > 1000
    _check(v, 1000)
