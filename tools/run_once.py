"""Profiling helper: pack + upload the cfg2 workload once and run the hot path a few times (for ncu)."""
import argparse
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "telomere-analyzer_b200"))
from nanotel_b200 import Scanner  # noqa: E402
from nanotel_b200.synth import synth_reads  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--reads", type=int, default=100000)
ap.add_argument("--runs", type=int, default=4)
ap.add_argument("--tvr", default=None)
ap.add_argument("--patterns", default="YYAGGG")
ap.add_argument("--no-jit", action="store_true")
ap.add_argument("--n-frac", type=float, default=None, help="fraction of reads that carry N (they take the 4-bit path)")
a = ap.parse_args()
buf, off, meta = synth_reads(a.reads, 20261020, **({} if a.n_frac is None else {"n_frac": a.n_frac}))
sc = Scanner(a.patterns, a.tvr, rc=True, jit=False if a.no_jit else None)
sc.pack_concat(buf, off)
sc.upload()
for _ in range(a.runs):
    sc.run()
t = sc.timings()
k = max(t["steps"], 1)
print("scan_ms %.4f triage_ms %.4f locate_ms %.4f (triage + locate) candidates %d jit %d" % (
    t["scan_ms"] / k, t["triage_ms"], t["locate_ms"] / k, t["candidates"], t["scan_is_jit"]))
