#!/usr/bin/env python
"""bench.py -- Gbases/s scanned (telomere calls bit-exact) on N B200s, with the HBM roofline and a CPU baseline.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--workload cfg2|cfg3|cfg4]
                    [--reads R] [--scaling weak|strong] [--no-subrecords] [--no-parity] [--no-cpu-baseline]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port P \
        bench.py --gpus N --steps K --warmup W

A "step" is one pass of the hot path (edge filter if enabled -> scan -> locate) over one batch of synthetic reads.
Workload (BASELINE.json configs[1], "cfg2"): 100 000 synthetic ONT-like reads per GPU (median 20 kb, ~2.3 Gbases,
10 % telomeric), --patterns YYAGGG --rc.  The global batch is made of blocks of 100 000 reads, block b generated from
seed 20261018 + 2 + 1000 b; rank r scans its blocks (weak scaling: one block per rank; --scaling strong: a fixed
total of 8 blocks = 800 000 reads cut over the ranks).  Reads are independent: no collective on the data path;
`value` = bases of all ranks / max-over-ranks device time.  Rank 0 gathers one digest per block of records after the
timed region (`block_digests`): block 0's digest is the same at every N, whoever scanned it.

  value      kernels only, packed reads resident in HBM (> 126 MB L2 per GPU, so every step streams from HBM)
  e2e        ntl_scan_batch(): pageable host ASCII in -> host results out (pack to pinned, H2D, kernels, D2H inside)
  roofline   scan kernel: algorithmic bytes (ceil(L*2/8) + T*n_win*2 + 64 per read) / CUDA-event time
  parity     every record and every window count of the timed batch against the oracle, outside the timed regions
  sub        (N = 1) compact records for the other BASELINE configurations: cfg3, cfg4 at S = 100 / 200 / 500, a
             10 000-read (--nrec-sized) batch
  cpu_baseline / --impl reference: the oracle (CPU restatement of NanoTel.R; R itself is not installable here)
             on all host cores over a bounded sample of the same workload.
"""
from __future__ import annotations

import argparse
import hashlib
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "telomere-analyzer_b200"))

import numpy as np  # noqa: E402

WORKLOADS = {
    # name: (patterns, tvr, rc, use_filter, right_edge, S, description)
    "cfg2": ("YYAGGG", None, True, False, False, 100,
             "synthetic ONT-like reads (median 20 kb, 10% telomeric), --patterns YYAGGG --rc"),
    "cfg3": ("YYAGGG", "TTGGG CCAGGG TCAGGG", True, False, False, 100,
             "cfg2 + --tvr_patterns 'TTGGG CCAGGG TCAGGG' (3 tracks)"),
    "cfg4": ("TTAGGG", None, False, True, True, 100,
             "synthetic reads, --patterns TTAGGG --use_filter --check_right_edge"),
}
SEED = 20261018
BLOCK_READS = 100000
STRONG_BLOCKS = 8
METRIC = "Gbases/s scanned (telomere calls bit-exact)"
DTYPE = "u32 bit-planes + f64 densities"


def algorithmic_bytes(lengths: np.ndarray, S: int, T: int) -> int:
    """per read ceil(L*2/8) + T*n_win*2 + 64 (2-bit reads)."""
    L = lengths.astype(np.int64)
    n0 = (L - 1) // S + 1
    last = 1 + (n0 - 1) * S
    n_win = n0 - (2 * (L - last) < S)
    return int(((L * 2 + 7) // 8).sum() + (T * n_win * 2).sum() + 64 * len(L))


class ClockSampler(threading.Thread):
    """nvidia-smi clocks / throttle reasons during the timed region (B200_PROFILING.md clocks line)."""

    def __init__(self, gpu_index: int):
        super().__init__(daemon=True)
        self.gpu = gpu_index
        self.rows = []
        self._stop = threading.Event()
        self.proc = None

    def run(self):
        q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
             "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
             "clocks_event_reasons.sw_power_cap")
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.gpu), "--query-gpu=" + q,
                                          "--format=csv,noheader,nounits", "-lms", "20"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            for line in self.proc.stdout:
                if self._stop.is_set():
                    break
                self.rows.append([x.strip() for x in line.split(",")])
        except Exception:
            pass

    def stop(self) -> dict:
        self._stop.set()
        if self.proc:
            self.proc.terminate()
        sm, mx, reasons = [], [], set()
        for r in self.rows:
            try:
                sm.append(float(r[0])); mx.append(float(r[1]))
                for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), r[3:7]):
                    if v.lower().startswith("active"):
                        reasons.add(name)
            except Exception:
                continue
        if not sm:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "samples": 0}
        return {"sm_mhz": float(np.median(sm)), "sm_max_mhz": float(max(mx)), "reasons": sorted(reasons),
                "samples": len(sm)}


def measured_peak_hbm():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        try:
            return float(json.load(open(p))["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
        except Exception:
            pass
    return 6650.0, "fallback (B200_PROFILING.md: 6.65 TB/s)"


def ncu_traffic(key: str):
    """dram bytes per scan-kernel launch from the committed ncu --set full capture, if one exists for this workload."""
    p = os.path.join(ROOT, "profiles", "scan_traffic.json")
    if os.path.exists(p):
        try:
            return json.load(open(p)).get(key)
        except Exception:
            return None
    return None


def cpu_baseline(buf, offsets, wl, n_sample: int, threads: int):
    """The oracle on `threads` host cores over the first n_sample reads of the workload.  TEST INFRASTRUCTURE used
    as the reported CPU baseline only."""
    from oracle import oracle as O
    patterns, tvr, rc, use_filter, right_edge, S, _ = wl
    n = min(n_sample, len(offsets) - 1)
    seqs = [buf[int(offsets[i]):int(offsets[i + 1])].tobytes() for i in range(n)]
    bases = int(offsets[n] - offsets[0])
    P = O.make_params(patterns, tvr, 0.6, S, right_edge)
    t0 = time.perf_counter()
    O.scan_batch(P, seqs, do_rc=rc, use_filter=use_filter, n_threads=threads, want_windows=False)
    dt = time.perf_counter() - t0
    return bases / dt / 1e9, bases, n, dt


def records_digest(res: np.ndarray) -> str:
    r = res.copy()
    r["win_offset"] = 0                          # internal: position inside the owning device's count planes
    return hashlib.sha256(r.tobytes()).hexdigest()[:16]


def device_steps(sc, steps: int):
    """K back-to-back passes on the resident batch; returns per-step kernel times (ms) from the library's events."""
    done = 0
    acc = {"filter": 0.0, "scan": 0.0, "triage": 0.0, "locate": 0.0, "launches": 0, "steps": 0}
    while done < steps:
        k = min(256, steps - done)
        for _ in range(k):
            sc.enqueue()
        sc.wait()
        tm = sc.timings()
        acc["filter"] += tm["filter_ms"]; acc["scan"] += tm["scan_ms"]; acc["triage"] += tm["triage_ms"]
        acc["locate"] += tm["locate_ms"] - tm["triage_ms"]; acc["launches"] += tm["kernel_launches"]
        acc["steps"] += tm["steps"]
        done += k
    n = max(acc["steps"], 1)
    return {k: acc[k] / n for k in ("filter", "scan", "triage", "locate")}, acc["launches"], tm


def sub_record(name, wl, buf, offsets, meta, steps, cores, do_parity, e2e_steps=0, traffic_key=None):
    """A compact record of another BASELINE configuration on rank 0's GPU (single process, no barriers)."""
    import torch
    from nanotel_b200 import Scanner
    patterns, tvr, rc, use_filter, right_edge, S, desc = wl
    T = 3 if tvr else 2
    bases = int(meta["bases"])
    with Scanner(patterns, tvr, 0.6, S, rc=rc, use_filter=use_filter, right_edge=right_edge) as sc:
        sc.pack_concat(buf, offsets)
        sc.upload()
        for _ in range(3):
            sc.run()
        stream = torch.cuda.ExternalStream(sc.stream)
        ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize()
        ev0.record(stream)
        for _ in range(steps):
            sc.enqueue()
        ev1.record(stream)
        sc.wait()
        torch.cuda.synchronize()
        ms = ev0.elapsed_time(ev1) / steps
        km, launches, tm = device_steps(sc, steps)
        res = sc.download().copy()
        scanned = (res["status"] & 2) == 0
        alg = algorithmic_bytes(meta["lengths"][scanned], S, T)
        peak, _ = measured_peak_hbm()
        rec = {"workload": name, "config": desc + ", subseq_length %d" % S, "reads": int(len(res)), "bases": bases,
               "value": bases / (ms * 1e-3) / 1e9, "unit": "Gbases/s", "ms_per_step": ms, "steps": steps,
               "kernel_ms": km, "scan_path": sc.scan_path,
               "roofline": {"kernel_ms": km["scan"], "algorithmic_bytes_per_launch": alg,
                            "achieved": alg / (km["scan"] * 1e-3) / 1e9 if km["scan"] > 0 else None,
                            "frac": alg / (km["scan"] * 1e-3) / 1e9 / peak if km["scan"] > 0 else None,
                            "traffic": ncu_traffic(traffic_key or name),
                            "reads_scanned": int(scanned.sum()), "bases_scanned": int(meta["lengths"][scanned].sum())},
               "reads_kept": int((res["status"] & 1).sum()), "candidates": int(tm["candidates"])}
        if e2e_steps:
            sc.scan_concat(buf, offsets, out="view")
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            for _ in range(e2e_steps):
                sc.scan_concat(buf, offsets, out="view")
            torch.cuda.synchronize()
            dt = (time.perf_counter() - t0) / e2e_steps
            rec["e2e"] = {"value": bases / dt / 1e9, "unit": "Gbases/s", "ms_per_step": dt * 1e3, "steps": e2e_steps}
        if do_parity:
            from oracle.compare import full_parity
            res_p = sc.download()
            rec["parity"] = full_parity(sc, res_p, (buf, offsets), patterns, tvr, 0.6, S, right_edge, rc, use_filter,
                                        n_threads=cores)
    return rec


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="cfg2", choices=sorted(WORKLOADS))
    ap.add_argument("--reads", type=int, default=BLOCK_READS, help="reads per block (one block per GPU in weak scaling)")
    ap.add_argument("--scaling", default="weak", choices=["weak", "strong"],
                    help="strong: a fixed batch of %d blocks cut over the ranks" % STRONG_BLOCKS)
    ap.add_argument("--e2e-steps", type=int, default=10)
    ap.add_argument("--cpu-sample", type=int, default=50000, help="reads of the CPU-baseline sample")
    ap.add_argument("--no-jit", action="store_true")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-parity", action="store_true", help="skip the full oracle comparison of the timed workload")
    ap.add_argument("--no-subrecords", action="store_true", help="skip the cfg3 / cfg4 / small-batch sub-records (N = 1)")
    args = ap.parse_args()

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    wl = WORKLOADS[args.workload]
    patterns, tvr, rc, use_filter, right_edge, S, desc = wl
    T = 3 if tvr else 2
    cores = os.cpu_count() or 1
    n_blocks_total = STRONG_BLOCKS if args.scaling == "strong" else world
    config = {"workload": "%s: %d reads/block, %d block(s)%s, %s" % (
                  args.workload, args.reads, n_blocks_total,
                  " in total over all ranks" if args.scaling == "strong" else " (one per GPU)", desc),
              "reads_per_block": args.reads, "blocks": n_blocks_total, "subseq_length": S, "min_density": 0.6,
              "tracks": T, "l2_policy": "inputs larger than L2 (packed reads >> 126 MB per GPU); no flush",
              "seed": SEED + 2}

    from nanotel_b200.synth import synth_reads

    # ---------------------------------------------------------------- reference arm: CPU only, rank 0 only
    if args.impl == "reference":
        if rank != 0:
            return
        n_gen = min(args.reads, max(args.cpu_sample, 1000))
        buf, offsets, meta = synth_reads(n_gen, SEED + 2)
        vals = []
        for it in range(args.warmup + args.steps):
            v, bases, n, dt = cpu_baseline(buf, offsets, wl, args.cpu_sample, cores)
            if it >= args.warmup:
                vals.append((v, dt))
        value = float(np.mean([v for v, _ in vals]))
        ms = float(np.mean([dt for _, dt in vals]) * 1e3)
        sample = "first %d reads (%d bases) of the %s workload per step" % (n, bases, args.workload)
        line = {"impl": "reference", "metric": METRIC, "value": value,
                "unit": "Gbases/s", "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
                "ms_per_step": ms, "higher_is_better": True, "scaling": args.scaling, "vs_baseline": None,
                "dtype": DTYPE, "data": "synthetic", "config": config,
                "cpu_baseline": {"value": value, "unit": "Gbases/s", "cores": cores, "kind": "port", "sample": sample,
                                 "note": "CPU restatement of NanoTel.R (oracle/), not the R/Biostrings path: R is "
                                         "not installable in this image"},
                "e2e": {"value": value, "unit": "Gbases/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
        print(json.dumps(line))
        return

    # ---------------------------------------------------------------- our arm
    # the ranks of one box share its host: every rank keeps its packer threads (and, by first touch, its buffers) on
    # its own slice of the cores, so that a rank's memory traffic stays on one socket
    host_threads = max(1, min(64, cores // max(world, 1)))
    if world > 1:
        try:
            avail = sorted(os.sched_getaffinity(0))
            per = max(1, len(avail) // world)
            os.sched_setaffinity(0, set(avail[local_rank * per:(local_rank + 1) * per]) or set(avail))
        except Exception:
            pass

    # stdout carries exactly one JSON line: everything else that libraries print there (NCCL's version banner, ...)
    # goes to stderr -- file descriptor 1 is pointed at stderr until the line is written
    sys.stdout.flush()
    real_stdout = os.dup(1)
    os.dup2(2, 1)

    import torch
    import torch.distributed as dist

    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: libnanotel_b200 has no CPU fallback")
    torch.cuda.set_device(local_rank)
    if world > 1:
        # NCCL writes its banner / debug lines to stdout by default; stdout carries the one JSON line
        os.environ.setdefault("NCCL_DEBUG_FILE", "/dev/stderr")
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))

    from nanotel_b200 import Scanner

    # this rank's blocks of the global batch
    if args.scaling == "strong":
        my_blocks = [b for b in range(STRONG_BLOCKS) if b * world // STRONG_BLOCKS == rank] if world <= STRONG_BLOCKS \
            else ([rank] if rank < STRONG_BLOCKS else [])
    else:
        my_blocks = [rank]
    parts = [synth_reads(args.reads, SEED + 2 + 1000 * b) for b in my_blocks]
    if parts:
        buf = np.concatenate([p[0] for p in parts]) if len(parts) > 1 else parts[0][0]
        lens = np.concatenate([p[2]["lengths"] for p in parts])
    else:
        buf, lens = np.zeros(1, np.uint8), np.zeros(0, np.int64)
    offsets = np.zeros(len(lens) + 1, np.int64)
    np.cumsum(lens, out=offsets[1:])
    meta = {"lengths": lens, "bases": int(offsets[-1])}
    del parts
    bases = int(meta["bases"])
    sc = Scanner(patterns, tvr, 0.6, S, rc=rc, use_filter=use_filter, right_edge=right_edge, device=local_rank,
                 jit=False if args.no_jit else None, host_threads=host_threads)
    if sc.note and rank == 0:
        print(sc.note, file=sys.stderr)

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # -- device-resident timing: pack + upload once, then K timed passes
    sc.pack_concat(buf, offsets)
    sc.upload()
    for _ in range(max(args.warmup, 3)):
        sc.run()
    stream = torch.cuda.ExternalStream(sc.stream, device=torch.device("cuda", local_rank))
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    sampler = ClockSampler(local_rank) if rank == 0 else None
    if sampler:
        sampler.start()
    # untimed pre-roll (~0.5 s of back-to-back passes) so that the clock samples are taken under load; the timed
    # region itself lasts only K x 0.3 ms
    t_pre = time.perf_counter()
    while time.perf_counter() - t_pre < 0.5:
        for _ in range(64):
            sc.enqueue()
        sc.wait()
    barrier()
    ev0.record(stream)
    done = 0
    while done < args.steps:
        k = min(256, args.steps - done)
        for _ in range(k):
            sc.enqueue()
        if done + k >= args.steps:
            ev1.record(stream)
        sc.wait()
        done += k
    barrier()
    dev_ms = ev0.elapsed_time(ev1)
    km, launches, tm = device_steps(sc, args.steps)          # the same passes again for the per-kernel split

    # -- end to end through the public call, host buffers in and out
    # results are read where the C ABI leaves them (the library's pinned host buffer, valid until the next batch)
    res = sc.scan_concat(buf, offsets, out="view")           # warm-up (allocations)
    barrier()
    t0 = time.perf_counter()
    for _ in range(args.e2e_steps):
        res = sc.scan_concat(buf, offsets, out="view")
    torch.cuda.synchronize()
    e2e_s = (time.perf_counter() - t0) / args.e2e_steps
    tm_e2e = sc.timings()
    res = res.copy()                                          # outside the timed region: later batches reuse the buffer
    n_keep = int((res["status"] & 1).sum())

    # -- what the host's memory system gives the packer: every rank streams its own ASCII input with its own packer
    #    threads at the same time (loads only: the packer's access pattern without arithmetic or stores)
    from nanotel_b200 import _lib as _ntl_lib
    barrier()
    host_read_gbs = float(_ntl_lib.load().ntl_host_read_gbs(buf.ctypes.data, int(offsets[-1]), host_threads, 3)) if bases else 0.0
    barrier()
    host_rw_gbs = float(_ntl_lib.load().ntl_host_read_gbs(buf.ctypes.data, int(offsets[-1]), host_threads, -3)) if bases else 0.0

    # -- the same, starting from the packed reads in pinned host memory (H2D + kernels + D2H per step): what the
    #    path costs once the ASCII -> 2-bit packing is taken out (extra information, not the headline)
    sc.pack_concat(buf, offsets)
    barrier()
    t0 = time.perf_counter()
    for _ in range(3):
        sc.upload()
        sc.run()
        sc.download(out="view")
    torch.cuda.synchronize()
    prepacked_s = (time.perf_counter() - t0) / 3
    clocks = sampler.stop() if sampler else None

    # -- parity verdict on the timed workload (outside every timed region; rank 0's blocks): every per-read record and
    #    every window count of the batch that is resident right now against the oracle (TEST INFRASTRUCTURE as checker)
    parity = None
    if rank == 0 and not args.no_parity:
        from oracle.compare import full_parity
        res_p = sc.download()
        parity = full_parity(sc, res_p, (buf, offsets), patterns, tvr, 0.6, S, right_edge, rc, use_filter,
                             n_threads=cores)

    # -- max over ranks; one digest per block, gathered on rank 0 (the host-side gather of the design, after the
    #    timed region; NCCL carries 8 bytes per block here, nothing on the data path)
    t = torch.tensor([dev_ms, e2e_s], dtype=torch.float64, device="cuda")
    b = torch.tensor([float(bases), host_read_gbs, bases / max(tm_e2e["pack_ms"], 1e-9) / 1e6, host_rw_gbs], dtype=torch.float64, device="cuda")
    dig = torch.zeros(max(n_blocks_total, 1), dtype=torch.int64, device="cuda")
    pos = 0
    for j, blk in enumerate(my_blocks):
        d = records_digest(res[pos:pos + args.reads])
        dig[blk] = int(d[:15], 16)
        pos += args.reads
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        dist.all_reduce(b, op=dist.ReduceOp.SUM)
        dist.all_reduce(dig, op=dist.ReduceOp.SUM)
    dev_ms_max, e2e_s_max = float(t[0]), float(t[1])
    total_bases = float(b[0])
    block_digests = ["%015x" % int(x) for x in dig.tolist()]

    if rank == 0:
        ms_per_step = dev_ms_max / args.steps
        value = total_bases / (ms_per_step * 1e-3) / 1e9
        # with --use_filter the scan kernel only streams the reads that passed the edge filter
        scanned = (res["status"] & 2) == 0
        alg = algorithmic_bytes(meta["lengths"][scanned], S, T)
        peak, peak_src = measured_peak_hbm()
        achieved = alg / (km["scan"] * 1e-3) / 1e9
        line = {
            "metric": METRIC, "value": value, "unit": "Gbases/s",
            "n_gpus": world, "steps": args.steps, "warmup": max(args.warmup, 3), "ms_per_step": ms_per_step,
            "higher_is_better": True, "scaling": args.scaling, "vs_baseline": None,
            "dtype": DTYPE, "data": "synthetic", "config": config,
            "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                         "traffic": ncu_traffic(args.workload),
                         "kernel": "ntl_scan_jit (span scan, %s)" % sc.scan_path if tm["scan_is_jit"] else "ntl_scan_generic_kernel",
                         "peak_source": peak_src, "algorithmic_bytes_per_launch": alg,
                         "bytes_per_base": alg / max(bases, 1), "kernel_ms": km["scan"],
                         "step_level": {"achieved": alg / (ms_per_step * 1e-3) / 1e9 if world == 1 else None,
                                        "frac": alg / (ms_per_step * 1e-3) / 1e9 / peak if world == 1 else None,
                                        "note": "the same bytes over the whole step (rank 0)"}},
            "kernel_ms": km,
            "e2e": {"value": total_bases / e2e_s_max / 1e9, "unit": "Gbases/s",
                    "h2d_bytes_per_step": int(tm_e2e["h2d_bytes"]), "d2h_bytes_per_step": int(tm_e2e["d2h_bytes"]),
                    "ms_per_step": e2e_s_max * 1e3, "steps": args.e2e_steps,
                    "breakdown_ms": {k: tm_e2e[k] for k in ("pack_ms", "h2d_ms", "filter_ms", "scan_ms", "locate_ms", "d2h_ms")},
                    "host_threads": host_threads, "host_cores": cores,
                    "host_memory": {"read_gbs_all_ranks": float(b[1]), "read_write_gbs_all_ranks": float(b[3]),
                                    "pack_gbs_all_ranks": float(b[2]),
                                    "note": "GB/s of ASCII input moved by the packer threads of all ranks at once: plain "
                                            "streaming reads; reads + a quarter-size non-temporal write (the packer's "
                                            "traffic without its arithmetic); the packer itself inside the e2e steps "
                                            "(where the DMA engine also reads 0.25 B per base)"},
                    "from_packed_pinned": {"value": bases / prepacked_s / 1e9, "unit": "Gbases/s (rank 0)",
                                           "ms_per_step": prepacked_s * 1e3},
                    "note": "pack (host, AVX2) and H2D overlap: h2d_ms spans first to last copy; per-rank h2d/d2h bytes"},
            "gpu_launches": int(round(launches / max(args.steps, 1) * args.steps)),
            "clocks": clocks, "telomeric_reads_found_rank0": n_keep, "bases_rank0": bases,
            "locate_candidates_rank0": int(tm["candidates"]), "scan_path": sc.scan_path,
            "block_digests": block_digests,
            "parity": parity,
        }
    sc.close()
    if rank == 0:
        # -- compact records of the other BASELINE configurations (single process only)
        if world == 1 and not args.no_subrecords and args.workload == "cfg2" and args.scaling == "weak":
            sub = []
            dp = not args.no_parity
            sub.append(sub_record("cfg3", WORKLOADS["cfg3"], buf, offsets, meta, 10, cores, dp))
            nb = min(10000, len(lens))
            small_meta = {"lengths": lens[:nb], "bases": int(offsets[nb])}
            sub.append(sub_record("cfg2_nrec10000", WORKLOADS["cfg2"], buf[:int(offsets[nb])], offsets[:nb + 1], small_meta,
                                  20, cores, False, e2e_steps=10))
            b4, o4, m4 = synth_reads(args.reads, SEED + 4, telomeric_frac=0.30)
            for S4 in (100, 200, 500):
                w4 = list(WORKLOADS["cfg4"]); w4[5] = S4
                sub.append(sub_record("cfg4_S%d" % S4, tuple(w4), b4, o4, m4, 10, cores, dp, e2e_steps=3 if S4 == 100 else 0,
                                      traffic_key="cfg4"))
            line["sub"] = sub
        if not args.no_cpu_baseline and world == 1:          # the CPU baseline leg runs at N = 1 only
            v, sb, sn, dt = cpu_baseline(buf, offsets, wl, args.cpu_sample, cores)
            line["cpu_baseline"] = {"value": v, "unit": "Gbases/s", "cores": cores, "kind": "port",
                                    "sample": "first %d reads (%d bases) of rank 0's batch, %.1f s" % (sn, sb, dt)}
        sys.stdout.flush()
        os.write(real_stdout, (json.dumps(line) + "\n").encode())
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
