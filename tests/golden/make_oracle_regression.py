"""Digest of the oracle's own results on seeded synthetic reads for a set of configurations -> a regression fixture.
This is NOT a pin on the reference (the configurations below exercise the parts of NanoTel.R that no artefact of the
reference covers: the < 100 bp edge fallback, the 18-bp re-match, the stricter re-run, IUPAC patterns, TVR, --rc, the
filter, other window sizes); it only makes sure that the oracle the GPU path is compared with does not change silently.

Run here:  python tests/golden/make_oracle_regression.py   (rewrites oracle_regression.json)"""
import hashlib
import json
import os
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "telomere-analyzer_b200"))

CONFIGS = [
    dict(patterns="TTAGGG"),
    dict(patterns="YYAGGG", rc=True),
    dict(patterns="YYAGGG", tvr="TTGGG CCAGGG TCAGGG", rc=True),
    dict(patterns="TTAGGG", right_edge=True, use_filter=True),
    dict(patterns="CCCTAA", rc=False, S=200),
    dict(patterns="TTAGGG TTGGG", S=64, min_density=0.3),
    dict(patterns="RRTCCC", tvr="CCCAA", rc=True, S=500),
]


def digest(cfg, seqs):
    from oracle import oracle as O
    P = O.make_params(cfg["patterns"], cfg.get("tvr"), cfg.get("min_density", 0.6), cfg.get("S", 100), cfg.get("right_edge", False))
    recs, passed, win_off, wc = O.scan_batch(P, seqs, do_rc=cfg.get("rc", False), use_filter=cfg.get("use_filter", False), n_threads=4)
    h = hashlib.sha256()
    h.update(recs.tobytes()); h.update(passed.tobytes()); h.update(win_off.tobytes()); h.update(wc[:int(win_off[-1])].tobytes())
    return h.hexdigest(), int((recs["keep"] != 0).sum())


def reads():
    from nanotel_b200.synth import as_list, synth_reads
    buf, off, meta = synth_reads(600, seed=20261018 + 77, median_len=5000, max_len=60000, telomeric_frac=0.35, n_frac=0.02)
    return as_list(buf, off)


def main():
    seqs = reads()
    out = {"reads": "synth_reads(600, seed=20261095, median_len=5000, max_len=60000, telomeric_frac=0.35, n_frac=0.02)", "configs": []}
    for cfg in CONFIGS:
        d, k = digest(cfg, seqs)
        out["configs"].append({"config": cfg, "sha256": d, "kept": k})
        print(cfg, k, d[:16])
    json.dump(out, open(os.path.join(HERE, "oracle_regression.json"), "w"), indent=1)


if __name__ == "__main__":
    main()
