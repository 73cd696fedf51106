"""GPU parity: every result of the CUDA path (called through the C ABI) equals the CPU oracle bit for bit."""
import numpy as np
import pytest

from helpers import compare_batch, oracle_batch

pytestmark = pytest.mark.gpu


def _run(seqs, patterns, tvr=None, min_density=0.6, S=100, right_edge=False, rc=False, use_filter=False, jit=None,
         label=""):
    from nanotel_b200 import Scanner
    P, recs, passed, win_off, wc = oracle_batch(seqs, patterns, tvr, min_density, S, right_edge, rc, use_filter)
    with Scanner(patterns, tvr, min_density, S, rc=rc, use_filter=use_filter, right_edge=right_edge, jit=jit,
                 debug_stages=True) as sc:
        res = sc.scan(seqs)
        tm = sc.timings()
        if jit is True:
            assert tm["scan_is_jit"] == 1
        compare_batch(sc, res, seqs, recs, passed, win_off, wc, check_stages=True, label=label)
    return res


CONFIGS = [
    dict(patterns="TTAGGG"),
    dict(patterns="YYAGGG"),
    dict(patterns="YYAGGG", tvr="TTGGG CCAGGG TCAGGG"),
    dict(patterns="TTAGGG", S=200),
    dict(patterns="TTAGGG", S=500),
    dict(patterns="TTAGGG TTGGG"),
    dict(patterns="TTAGGG", right_edge=True),
    dict(patterns="CCCTAA", rc=True),
    dict(patterns="RRTCCC", rc=True, tvr="CCCAA"),
    dict(patterns="TTAGGG", min_density=0.3, S=64),
]


@pytest.mark.parametrize("jit", [False, True])
@pytest.mark.parametrize("cfg", CONFIGS, ids=lambda c: "-".join("%s=%s" % kv for kv in c.items()).replace(" ", "+"))
def test_example_reads(example_reads, cfg, jit):
    seqs = [s for _, s in example_reads]
    _run(seqs, jit=jit, label=str(cfg), **cfg)


@pytest.mark.parametrize("jit", [False, True])
def test_synthetic_cfg2_cfg3(jit):
    from nanotel_b200.synth import as_list, synth_reads
    buf, off, meta = synth_reads(400, seed=20261018 + 2, median_len=6000, max_len=60000, telomeric_frac=0.3)
    seqs = as_list(buf, off)
    _run(seqs, "YYAGGG", rc=True, jit=jit, label="cfg2")
    _run(seqs, "YYAGGG", tvr="TTGGG CCAGGG TCAGGG", rc=True, jit=jit, label="cfg3")
    _run(seqs, "TTAGGG", right_edge=True, jit=jit, label="fwd right edge")


@pytest.mark.parametrize("jit", [False, True])
def test_iupac_reads_take_the_4bit_path(jit):
    """Reads with N: four nibble planes, scanned by ntl_scan_kernel<4> (jit=False) or by the NVRTC build ntl_scan_jit4."""
    from nanotel_b200 import READ_IUPAC
    from nanotel_b200.synth import as_list, synth_reads
    buf, off, meta = synth_reads(120, seed=7, median_len=3000, max_len=20000, telomeric_frac=0.5, n_frac=0.5)
    seqs = as_list(buf, off)
    res = _run(seqs, "YYAGGG", rc=True, jit=jit, label="N reads")
    assert (res["status"] & READ_IUPAC).any()
    _run(seqs, "TTAGGG", tvr="TTGGG", jit=jit, label="N reads fixed")
    _run(seqs, "TTAGGG CCCTAA", tvr="TTGGG CCAGGG TCAGGG", jit=jit, label="N reads, two lengths")


def test_short_and_ragged_reads():
    rng = np.random.default_rng(5)
    seqs = []
    for L in list(range(1, 140)) + [149, 150, 151, 199, 200, 201, 249, 250, 251, 4095, 4096, 4097, 8191, 8192, 8193]:
        unit = b"TTAGGG" if L % 2 else b"CCCTAA"
        s = bytearray((unit * (L // 6 + 2))[:L])
        for _ in range(L // 40):
            s[int(rng.integers(0, L))] = b"ACGT"[int(rng.integers(0, 4))]
        seqs.append(bytes(s))
    for L in (31, 32, 33, 63, 64, 65, 127, 128, 129, 1000, 5000):
        seqs.append(bytes(rng.choice(np.frombuffer(b"ACGT", np.uint8), L)))
    _run(seqs, "TTAGGG", label="short fwd")
    _run(seqs, "CCCTAA", label="short C")
    _run(seqs, "TTAGGG", S=20, min_density=0.5, label="short S=20")
    _run(seqs, "AAAAAA TTTTT", label="self-overlapping")


def test_filter_matches_oracle():
    from nanotel_b200.synth import as_list, synth_reads
    buf, off, meta = synth_reads(300, seed=11, median_len=4000, min_len=500, max_len=30000, telomeric_frac=0.5)
    seqs = as_list(buf, off)
    _run(seqs, "TTAGGG", use_filter=True, right_edge=True, label="filter right")
    _run(seqs, "CCCTAA", use_filter=True, right_edge=False, label="filter left")
    _run(seqs, "YYAGGG", use_filter=True, right_edge=False, rc=True, label="filter rc")
