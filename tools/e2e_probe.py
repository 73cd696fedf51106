"""Where does an end-to-end ntl_scan_batch call spend its time? (host wall clock vs the library's own timings)"""
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "telomere-analyzer_b200"))
import numpy as np  # noqa: E402
from nanotel_b200 import Scanner  # noqa: E402
from nanotel_b200.synth import synth_reads  # noqa: E402

buf, off, meta = synth_reads(100000, 20261020)
sc = Scanner("YYAGGG", rc=True)
sc.scan_concat(buf, off)
for _ in range(3):
    t0 = time.perf_counter()
    b, ptrs, lens = sc._marshal_concat(buf, off)
    t1 = time.perf_counter()
    import ctypes as C
    out = C.c_void_p()
    sc._check(sc._L.ntl_scan_batch(sc._h, ptrs.ctypes.data, lens.ctypes.data, len(lens), C.byref(out)))
    t2 = time.perf_counter()
    res = sc._results_view(out, len(lens)).copy()
    t3 = time.perf_counter()
    tm = sc.timings()
    print("marshal %.2f  call %.2f  copy %.2f | lib total %.2f pack %.2f h2d %.2f scan %.3f locate %.3f d2h %.2f" % (
        (t1 - t0) * 1e3, (t2 - t1) * 1e3, (t3 - t2) * 1e3, tm["total_ms"], tm["pack_ms"], tm["h2d_ms"], tm["scan_ms"],
        tm["locate_ms"], tm["d2h_ms"]))
