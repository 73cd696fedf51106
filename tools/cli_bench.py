"""Wall clock of the R-free command line on synthetic gzip FASTQ files (cfg2-like reads):  cli_bench.py [n_reads] [n_files]"""
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
k = int(sys.argv[2]) if len(sys.argv) > 2 else 1
fq = os.path.join(d, "in")                              # -i takes a file or a directory of files
os.makedirs(fq)
for j in range(k):
    with gzip.open(os.path.join(fq, "part%03d.fastq.gz" % j), "wb", compresslevel=1) as f:
        for i in range(j * n // k, (j + 1) * n // k):
            f.write(b"@read%08d\n" % i + seqs[i] + b"\n+\n" + b"I" * len(seqs[i]) + b"\n")
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
