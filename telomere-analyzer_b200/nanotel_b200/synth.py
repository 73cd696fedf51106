"""Deterministic synthetic ONT-like reads for the BASELINE.json configurations.

The generator is ours (the reference ships only Example/sample.fasta).  Reads are returned as ONE contiguous ASCII
buffer plus offsets, which is what ntl_scan_batch_concat() takes; `as_list` gives a list of bytes for small cases.

  length        lognormal(ln median, sigma) clipped to [min_len, max_len]
  background    iid A/C/G/T with P = (0.295, 0.205, 0.205, 0.295)
  telomeric     `telomeric_frac` of the reads: half end in (TTAGGG)n (3' G-strand), half begin with (CCCTAA)n;
                tract ~ lognormal(ln 6000, 0.6) clipped to [200, 0.8 L], random phase, 0-80 nt adapter-like random
                letters outside the tract, 30 % with a 50-400 nt variant-repeat zone at the proximal boundary,
                ONT-like errors inside the tract (2 % substitutions, 1 % insertions, 1 % deletions)
  interstitial  0.5 % of the other reads carry a 60-300 nt telomeric block somewhere inside
  N             `n_frac` of the reads get 1-20 'N' (forces the 4-bit path; 0 for the throughput configurations)
"""
from __future__ import annotations

from typing import List, Tuple

import numpy as np

_ACGT = np.frombuffer(b"ACGT", dtype=np.uint8)
_COMP = np.zeros(256, np.uint8)
for _a, _b in zip(b"ACGTN", b"TGCAN"):
    _COMP[_a] = _b


def _background(rng: np.random.Generator, n: int) -> np.ndarray:
    """iid letters with P(A, C, G, T) = (0.295, 0.205, 0.205, 0.295), generated in slabs to bound memory."""
    out = np.empty(n, np.uint8)
    lut = np.empty(256, np.uint8)
    edges = np.round(np.cumsum([0.295, 0.205, 0.205, 0.295]) * 256).astype(int)
    lo = 0
    for k, hi in enumerate(edges):
        lut[lo:hi] = _ACGT[k]
        lo = hi
    slab = 1 << 26
    for s in range(0, n, slab):
        e = min(n, s + slab)
        out[s:e] = lut[rng.integers(0, 256, e - s, dtype=np.uint8)]
    return out


def _repeat(unit: bytes, n: int, phase: int) -> np.ndarray:
    u = np.frombuffer(unit, np.uint8)
    reps = (n + phase) // len(u) + 2
    return np.tile(u, reps)[phase:phase + n].copy()


def _with_errors(rng: np.random.Generator, seq: np.ndarray, sub=0.02, ins=0.01, dele=0.01) -> np.ndarray:
    n = len(seq)
    keep = rng.random(n) >= dele
    seq = seq[keep]
    n = len(seq)
    m = rng.random(n) < sub
    seq = seq.copy()
    seq[m] = _ACGT[rng.integers(0, 4, int(m.sum()))]
    ip = np.nonzero(rng.random(n) < ins)[0]
    if len(ip):
        seq = np.insert(seq, ip, _ACGT[rng.integers(0, 4, len(ip))])
    return seq


def _tract(rng: np.random.Generator, length: int, g_strand: bool, tvr: bool) -> np.ndarray:
    """A telomeric tract of exactly `length` letters on the G strand (TTAGGG...) or its reverse complement."""
    raw = _repeat(b"TTAGGG", int(length * 1.05) + 16, int(rng.integers(0, 6)))
    if tvr:
        zone = int(rng.integers(50, 401))
        units = [b"TTGGG", b"TCAGGG", b"CCAGGG", b"TTAGGG"]
        parts, tot = [], 0
        while tot < zone:
            u = units[int(rng.integers(0, 4))]
            parts.append(np.frombuffer(u, np.uint8))
            tot += len(u)
        z = np.concatenate(parts)[:min(zone, len(raw))]
        raw[:len(z)] = z                     # proximal (centromere side) boundary of a G-strand tract is its start
    seq = _with_errors(rng, raw)
    if len(seq) < length:
        seq = np.concatenate([seq, _repeat(b"TTAGGG", length - len(seq), 0)])
    seq = seq[:length]
    if not g_strand:
        seq = _COMP[seq[::-1]]
    return seq


def synth_reads(n_reads: int, seed: int, telomeric_frac: float = 0.10, median_len: float = 20000.0,
                sigma: float = 0.55, min_len: int = 1000, max_len: int = 250000, n_frac: float = 0.0,
                interstitial_frac: float = 0.005) -> Tuple[np.ndarray, np.ndarray, dict]:
    rng = np.random.default_rng(seed)
    lens = np.exp(rng.normal(np.log(median_len), sigma, n_reads))
    lens = np.clip(lens, min_len, max_len).astype(np.int64)
    offsets = np.zeros(n_reads + 1, np.int64)
    np.cumsum(lens, out=offsets[1:])
    buf = _background(rng, int(offsets[-1]))
    is_telo = rng.random(n_reads) < telomeric_frac
    kinds = np.zeros(n_reads, np.int8)          # 0 none, 1 G-strand at 3' end, 2 C-strand at 5' end, 3 interstitial
    for i in np.nonzero(is_telo)[0]:
        L = int(lens[i])
        t = int(min(max(np.exp(rng.normal(np.log(6000.0), 0.6)), 200.0), max(0.8 * L, 1.0)))
        g = bool(rng.random() < 0.5)
        tvr = bool(rng.random() < 0.3)
        adapter = int(rng.integers(0, 81))
        if adapter + t > L:
            adapter = 0
        tr = _tract(rng, t, g, tvr)
        s = int(offsets[i])
        if g:
            buf[s + L - adapter - t:s + L - adapter] = tr
            kinds[i] = 1
        else:
            buf[s + adapter:s + adapter + t] = tr
            kinds[i] = 2
    others = np.nonzero(~is_telo)[0]
    inter = others[rng.random(len(others)) < interstitial_frac]
    for i in inter:
        L = int(lens[i])
        t = int(rng.integers(60, 301))
        if t + 200 >= L:
            continue
        pos = int(rng.integers(100, L - t - 100))
        buf[int(offsets[i]) + pos:int(offsets[i]) + pos + t] = _tract(rng, t, bool(rng.random() < 0.5), False)
        kinds[i] = 3
    if n_frac > 0:
        for i in np.nonzero(rng.random(n_reads) < n_frac)[0]:
            k = int(rng.integers(1, 21))
            pos = rng.integers(0, int(lens[i]), k)
            buf[int(offsets[i]) + pos] = ord("N")
    return buf, offsets, {"lengths": lens, "kinds": kinds, "bases": int(offsets[-1])}


def as_list(buf: np.ndarray, offsets: np.ndarray) -> List[bytes]:
    return [buf[int(offsets[i]):int(offsets[i + 1])].tobytes() for i in range(len(offsets) - 1)]
