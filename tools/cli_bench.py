"""File-to-summary wall clock of the R-free command line (python -m nanotel_b200) on synthetic gzip FASTQ files, the
shape of BASELINE.json configs[3]: --use_filter --check_right_edge, --nrec batching, subseq_length sweep.

    python tools/cli_bench.py [--reads 100000] [--files 16] [--nrec 10000 100000] [--S 100 200 500] [--devices all]

One JSON line per (nrec, S, devices) run: seconds from process start to summary.csv on disk (CUDA context, reader,
scan, per-read FASTA.gz / density-vector files included), reads, bases, telomeric reads written."""
import argparse
import gzip
import json
import os
import subprocess
import sys
import tempfile
import time
from concurrent.futures import ThreadPoolExecutor

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "telomere-analyzer_b200"))
from nanotel_b200.synth import synth_reads  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--reads", type=int, default=100000)
ap.add_argument("--files", type=int, default=16)
ap.add_argument("--nrec", type=int, nargs="+", default=[10000, 100000])
ap.add_argument("--S", type=int, nargs="+", default=[100])
ap.add_argument("--devices", default=None)
ap.add_argument("--no-filter", action="store_true")
a = ap.parse_args()

d = tempfile.mkdtemp(prefix="ntl_cli_")
fq = os.path.join(d, "in")                              # -i takes a file or a directory of files
os.makedirs(fq)
# the input is generated in blocks of at most 100 000 reads (own seed each, a.files files per block) so that a
# 1 000 000-read run (BASELINE.json configs[3]) never holds more than one block of ASCII in memory
BLOCK = 100000
t0 = time.perf_counter()
total_bases = 0
for blk in range((a.reads + BLOCK - 1) // BLOCK):
    nb = min(BLOCK, a.reads - blk * BLOCK)
    buf, off, meta = synth_reads(nb, 20261018 + 4 + 1000 * blk, telomeric_frac=0.30)
    total_bases += int(meta["bases"])

    def write_part(j, blk=blk, nb=nb, buf=buf, off=off):
        with gzip.open(os.path.join(fq, "part%03d_%03d.fastq.gz" % (blk, j)), "wb", compresslevel=1) as f:
            for i in range(j * nb // a.files, (j + 1) * nb // a.files):
                s = buf[int(off[i]):int(off[i + 1])].tobytes()
                f.write(b"@read%08d\n" % (blk * BLOCK + i) + s + b"\n+\n" + b"I" * len(s) + b"\n")

    with ThreadPoolExecutor(max_workers=min(a.files, os.cpu_count() or 1)) as pool:
        list(pool.map(write_part, range(a.files)))
    del buf, off
meta = {"bases": total_bases}
gz_bytes = sum(os.path.getsize(os.path.join(fq, f)) for f in os.listdir(fq))
print(json.dumps({"generated": len(os.listdir(fq)), "gz_bytes": gz_bytes, "seconds": round(time.perf_counter() - t0, 1)}), flush=True)
env = dict(os.environ, PYTHONPATH=os.path.join(ROOT, "telomere-analyzer_b200"))
run = 0
for S in a.S:
    for nrec in a.nrec:
        out = os.path.join(d, "out%d" % run); run += 1
        cmd = [sys.executable, "-m", "nanotel_b200", "-i", fq, "--save_path", out, "--patterns", "TTAGGG",
               "--subseq_length", str(S), "--nrec", str(nrec)]
        if not a.no_filter:
            cmd += ["--use_filter", "--check_right_edge"]
        if a.devices:
            cmd += ["--devices", a.devices]
        t0 = time.perf_counter()
        subprocess.run(cmd, env=env, check=True, stdout=subprocess.DEVNULL)
        dt = time.perf_counter() - t0
        kept = len(os.listdir(os.path.join(out, "reads"))) if os.path.isdir(os.path.join(out, "reads")) else 0
        print(json.dumps({"nrec": nrec, "subseq_length": S, "devices": a.devices or "0", "filter": not a.no_filter,
                          "seconds": round(dt, 2), "reads": a.reads, "gbases": round(meta["bases"] / 1e9, 3),
                          "telomeric_reads_written": kept, "mbases_per_s": round(meta["bases"] / dt / 1e6, 1),
                          "host_cores": os.cpu_count()}), flush=True)
