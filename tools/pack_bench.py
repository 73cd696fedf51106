"""Host packer throughput vs thread count (ntl_batch_pack only; no kernels)."""
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "telomere-analyzer_b200"))
from nanotel_b200 import Scanner  # noqa: E402
from nanotel_b200.synth import synth_reads  # noqa: E402

MEDIAN = float(os.environ.get("PACK_MEDIAN", "20000"))      # PACK_MEDIAN=200000 with a tenth of the reads: per-read overhead
buf, off, meta = synth_reads(int(sys.argv[1]) if len(sys.argv) > 1 else 100000, 20261020, median_len=MEDIAN,
                             max_len=max(250000, int(MEDIAN * 10)))
RC = os.environ.get("PACK_RC", "1") == "1"
for nt in [int(x) for x in os.environ.get("PACK_THREADS", "1,4,16").split(",")]:
    sc = Scanner("YYAGGG", rc=RC, host_threads=nt)
    best = 1e9
    for _ in range(4):
        t0 = time.perf_counter()
        sc.pack_concat(buf, off)
        best = min(best, time.perf_counter() - t0)
    tm = sc.timings()
    print("threads %2d  wall %.1f ms  pack_ms %.1f  -> %.1f GB/s ASCII" % (nt, best * 1e3, tm["pack_ms"], meta["bases"] / best / 1e9))
    sc.close()
