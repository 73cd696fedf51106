"""GPU: seeded random differential test -- random pattern sets (IUPAC letters, 1..18 nt, lists and TVR lists), random
--subseq_length / --min_density / --rc / --check_right_edge / --use_filter, on small reads with planted repeats,
mutated repeats, repeats at both read ends, short reads and reads with IUPAC letters.  Every record, window count and
intermediate stage must equal the oracle bit for bit, for the NVRTC-specialised and the runtime-pattern kernel."""
import numpy as np
import pytest

from helpers import compare_batch, oracle_batch

pytestmark = pytest.mark.gpu

LETTERS = np.frombuffer(b"ACGT", np.uint8)
IUPAC = "ACGTRYSWKMBDHVN"


def _rand_pattern(rng, lo=3, hi=10, p_amb=0.15):
    m = int(rng.integers(lo, hi + 1))
    return "".join(IUPAC[int(rng.integers(4, len(IUPAC)))] if rng.random() < p_amb else "ACGT"[int(rng.integers(0, 4))]
                   for _ in range(m))


def _instance(pat, rng):
    """a concrete ACGT word matching the (possibly ambiguous) pattern"""
    code = {"A": "A", "C": "C", "G": "G", "T": "T", "R": "AG", "Y": "CT", "S": "CG", "W": "AT", "K": "GT", "M": "AC",
            "B": "CGT", "D": "AGT", "H": "ACT", "V": "ACG", "N": "ACGT"}
    return "".join(code[c][int(rng.integers(0, len(code[c])))] for c in pat).encode()


def _reads(rng, pats, n=60):
    out = []
    for i in range(n):
        L = int(rng.choice([int(rng.integers(1, 60)), int(rng.integers(60, 400)), int(rng.integers(400, 3000)),
                            int(rng.integers(3000, 9000))]))
        s = bytearray(rng.choice(LETTERS, L).tobytes())
        kind = int(rng.integers(0, 6))
        if kind >= 2 and L > 30:
            unit = _instance(pats[int(rng.integers(0, len(pats)))], rng)
            tl = int(rng.integers(10, max(11, int(L * rng.uniform(0.1, 1.0)))))
            rep = bytearray((unit * (tl // len(unit) + 2))[:tl])
            for _ in range(int(tl * rng.uniform(0, 0.08))):                 # mutations
                rep[int(rng.integers(0, tl))] = LETTERS[int(rng.integers(0, 4))]
            pos = {2: 0, 3: L - tl, 4: int(rng.integers(0, L - tl + 1)), 5: 0}[kind]
            s[pos:pos + tl] = rep
            if kind == 5 and L > 2 * tl:                                      # repeats at both ends
                s[L - tl:] = rep
        if rng.random() < 0.15 and L > 5:                                     # IUPAC letters in the read
            for _ in range(int(rng.integers(1, 4))):
                s[int(rng.integers(0, L))] = ord("NRYKMSW"[int(rng.integers(0, 7))])
        out.append(bytes(s))
    return out


@pytest.mark.parametrize("seed", range(24))
def test_random_configuration(seed):
    from nanotel_b200 import Scanner
    rng = np.random.default_rng(1000 + seed)
    n_pat = int(rng.choice([1, 1, 1, 2, 3]))
    pats = [_rand_pattern(rng, 3, 18 if seed % 6 == 0 else 9) for _ in range(n_pat)]
    if seed % 4 == 1:
        pats[0] = "TTAGGG"
    tvr = [_rand_pattern(rng, 3, 8, 0.1) for _ in range(int(rng.integers(1, 4)))] if rng.random() < 0.4 else None
    S = int(rng.choice([7, 20, 33, 64, 100, 100, 128, 200, 500]))
    md = float(rng.choice([0.1, 0.3, 0.5, 0.6, 0.6, 0.75, 0.9]))
    rc, right, filt = bool(rng.random() < 0.4), bool(rng.random() < 0.5), bool(rng.random() < 0.25)
    seqs = _reads(rng, pats)
    P, recs, passed, win_off, wc = oracle_batch(seqs, pats, tvr, md, S, right, rc, filt)
    label = "seed %d pats %s tvr %s S %d md %s rc %s right %s filter %s" % (seed, pats, tvr, S, md, rc, right, filt)
    for jit in (True, False):
        with Scanner(pats, tvr, md, S, rc=rc, use_filter=filt, right_edge=right, jit=jit, debug_stages=True) as sc:
            res = sc.scan(seqs)
            compare_batch(sc, res, seqs, recs, passed, win_off, wc, check_stages=True, label=label + " jit=%s" % jit)
