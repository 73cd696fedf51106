"""Throughput of the native FASTQ(.gz) reader alone on a list of files:  reader_bench.py [n_reads] [n_files]"""
import gzip
import os
import sys
import tempfile
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "telomere-analyzer_b200"))
from nanotel_b200.nanotel import NativeReader  # noqa: E402
from nanotel_b200.synth import as_list, synth_reads  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 20000
k = int(sys.argv[2]) if len(sys.argv) > 2 else 16
buf, off, meta = synth_reads(n, 20261022)
seqs = as_list(buf, off)
d = tempfile.mkdtemp(prefix="ntl_rd_")
paths = []
for j in range(k):
    p = os.path.join(d, "part%03d.fastq.gz" % j)
    with gzip.open(p, "wb", compresslevel=1) as f:
        for i in range(j * n // k, (j + 1) * n // k):
            f.write(b"@read%08d\n" % i + seqs[i] + b"\n+\n" + b"I" * len(seqs[i]) + b"\n")
    paths.append(p)
size = sum(os.path.getsize(p) for p in paths)
for ahead in (1, 2, 4, 8, 16):
    os.environ["NTL_READER_FILES"] = str(ahead)
    t0 = time.perf_counter()
    got = bases = 0
    for names, b, so in NativeReader(paths, "fastq", 10000):
        got += len(names); bases += int(so[-1])
    dt = time.perf_counter() - t0
    assert got == n and bases == meta["bases"]
    print("files inflated side by side %2d: %.2f s  %.0f Mbases/s  (%.0f MB/s of .gz)" % (ahead, dt, bases / dt / 1e6, size / dt / 1e6))
