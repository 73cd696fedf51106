"""Wall clock of the R-free command line on a synthetic gzip FASTQ (cfg2-like reads):  cli_bench.py [n_reads]"""
import gzip
import os
import subprocess
import sys
import tempfile
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "telomere-analyzer_b200"))
from nanotel_b200.synth import as_list, synth_reads  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 20000
buf, off, meta = synth_reads(n, 20261022)
seqs = as_list(buf, off)
d = tempfile.mkdtemp(prefix="ntl_cli_")
fq = os.path.join(d, "reads.fastq.gz")
with gzip.open(fq, "wb", compresslevel=1) as f:
    for i, s in enumerate(seqs):
        f.write(b"@read%08d\n" % i + s + b"\n+\n" + b"I" * len(s) + b"\n")
env = dict(os.environ, PYTHONPATH=os.path.join(ROOT, "telomere-analyzer_b200"))
for rep in range(2):                                   # the second run finds the NVRTC cubin in the cache
    out = os.path.join(d, "out%d" % rep)
    t0 = time.perf_counter()
    subprocess.run([sys.executable, "-m", "nanotel_b200", "-i", fq, "--save_path", out, "--patterns", "YYAGGG", "--rc"],
                   env=env, check=True, stdout=subprocess.DEVNULL)
    dt = time.perf_counter() - t0
    kept = len(os.listdir(os.path.join(out, "reads"))) if os.path.isdir(os.path.join(out, "reads")) else 0
    print("run %d: %.2f s for %d reads, %.3f Gbases (%d telomeric reads written): %.1f Mbases/s" % (
        rep, dt, n, meta["bases"] / 1e9, kept, meta["bases"] / dt / 1e6))
