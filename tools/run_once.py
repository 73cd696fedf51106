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
ap.add_argument("--use-filter", action="store_true")
ap.add_argument("--right-edge", action="store_true")
ap.add_argument("--no-rc", action="store_true")
ap.add_argument("--subseq", type=int, default=100)
ap.add_argument("--telomeric-frac", type=float, default=None)
ap.add_argument("--seed", type=int, default=20261020)
ap.add_argument("--n-frac", type=float, default=None, help="fraction of reads that carry N (they take the 4-bit path)")
ap.add_argument("--k3clock", action="store_true", help="print the locate kernel's per-phase cycles (a -DNTL_K3_CLOCK build)")
a = ap.parse_args()
kw = {}
if a.n_frac is not None:
    kw["n_frac"] = a.n_frac
if a.telomeric_frac is not None:
    kw["telomeric_frac"] = a.telomeric_frac
buf, off, meta = synth_reads(a.reads, a.seed, **kw)
sc = Scanner(a.patterns, a.tvr, 0.6, a.subseq, rc=not a.no_rc, use_filter=a.use_filter, right_edge=a.right_edge,
             jit=False if a.no_jit else None)
sc.pack_concat(buf, off)
sc.upload()
for _ in range(a.runs):
    sc.run()
t = sc.timings()
k = max(t["steps"], 1)
print("filter_ms %.4f scan_ms %.4f triage_ms %.4f locate_ms %.4f (triage + locate) candidates %d jit %d path %s" % (
    t["filter_ms"] / k, t["scan_ms"] / k, t["triage_ms"] / k, t["locate_ms"] / k, t["candidates"], t["scan_is_jit"], sc.scan_path))
if a.k3clock:
    import ctypes as C
    from nanotel_b200 import _lib
    L = _lib.load()
    out = (C.c_ulonglong * 48)()
    L.ntl_debug_k3_clock.argtypes = [C.POINTER(C.c_ulonglong)]
    if L.ntl_debug_k3_clock(out) > 0:
        names = ["setup", "coarse", "coarse density", "re-run", "acc start", "acc end", "edge/stage", "search right", "search left", "final density"]
        for ph in list(range(10)) + [15]:
            s_, m_, n_ = out[3 * ph], out[3 * ph + 1], out[3 * ph + 2]
            if n_:
                print("phase %2d %-15s mean %8.0f max %8d cycles  (%d warp passes)" % (ph, names[ph] if ph < 10 else "whole item group", s_ / n_, m_, n_))
