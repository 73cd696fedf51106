"""SHA-256 of the FASTA files the reference wrote for its own example (Example/Example_output/reads/<Serial>.fasta,
writeXStringSet: header line, 80 letters per line) -> a tiny fixture that pins the writer of the R-free driver.

Run here (needs /root/reference):  python tests/golden/make_reads_fasta_sha.py"""
import hashlib
import json
import os

DIR = "/root/reference/Example/Example_output/reads"
OUT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "example_reads_fasta_sha256.json")
res = {"source": "Example/Example_output/reads/<Serial>.fasta", "sha256": {}, "bytes": {}}
for k in (1, 2, 3, 4):
    raw = open(os.path.join(DIR, "%d.fasta" % k), "rb").read()
    res["sha256"][str(k)] = hashlib.sha256(raw).hexdigest()
    res["bytes"][str(k)] = len(raw)
json.dump(res, open(OUT, "w"), indent=1)
print(res)
