"""cProfile of the command-line driver on a synthetic gzip FASTQ: where the wall clock goes."""
import cProfile
import gzip
import os
import pstats
import sys
import tempfile

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "telomere-analyzer_b200"))
from nanotel_b200 import nanotel  # noqa: E402
from nanotel_b200.synth import as_list, synth_reads  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 20000
buf, off, meta = synth_reads(n, 20261022)
d = tempfile.mkdtemp(prefix="ntl_cli_")
fq = os.path.join(d, "reads.fastq.gz")
with gzip.open(fq, "wb", compresslevel=1) as f:
    for i, s in enumerate(as_list(buf, off)):
        f.write(b"@read%08d\n" % i + s + b"\n+\n" + b"I" * len(s) + b"\n")
argv = ["-i", fq, "--save_path", os.path.join(d, "out"), "--patterns", "YYAGGG", "--rc"]
cProfile.run("nanotel.main(argv)", os.path.join(d, "prof"))
pstats.Stats(os.path.join(d, "prof")).sort_stats("cumulative").print_stats(22)
