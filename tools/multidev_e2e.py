"""End-to-end throughput of ONE process that shards every batch over several GPUs (ntl_params.device_ids) against the
same batch on one GPU:   python tools/multidev_e2e.py [--reads 200000] [--steps 5]
Prints one JSON line per device count: Gbases/s through ntl_scan_batch_concat (host ASCII in, host records out), the
device-resident step, and whether the gathered records equal the single-device ones."""
import argparse
import hashlib
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "telomere-analyzer_b200"))
from nanotel_b200 import Scanner  # noqa: E402
from nanotel_b200.synth import synth_reads  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--reads", type=int, default=200000)
ap.add_argument("--steps", type=int, default=5)
a = ap.parse_args()
import torch  # noqa: E402

n_gpu = torch.cuda.device_count()
buf, off, meta = synth_reads(a.reads, 20261018 + 5)
ref = None
for k in sorted({1, 2, 4, 8, n_gpu}):
    if k > n_gpu:
        continue
    with Scanner("YYAGGG", None, rc=True, devices=list(range(k))) as sc:
        res = sc.scan_concat(buf, off, out="view")
        t0 = time.perf_counter()
        for _ in range(a.steps):
            res = sc.scan_concat(buf, off, out="view")
        dt = (time.perf_counter() - t0) / a.steps
        tm = sc.timings()
        r = res.copy(); r["win_offset"] = 0
        dig = hashlib.sha256(r.tobytes()).hexdigest()[:16]
        ref = ref or dig
        sc.pack_concat(buf, off); sc.upload()
        for _ in range(3):
            sc.run()
        t0 = time.perf_counter()
        for _ in range(20):
            sc.enqueue()
        sc.wait()
        dev = (time.perf_counter() - t0) / 20
        print(json.dumps({"devices": k, "reads": a.reads, "bases": int(meta["bases"]), "e2e_gbases_s": meta["bases"] / dt / 1e9,
                          "e2e_ms": dt * 1e3, "pack_ms": tm["pack_ms"], "device_step_ms_wall": dev * 1e3,
                          "device_gbases_s": meta["bases"] / dev / 1e9, "records_equal_single_device": dig == ref,
                          "host_cores": os.cpu_count()}))
