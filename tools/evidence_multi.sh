#!/bin/bash
# Multi-GPU evidence (run with `gpurun --gpus N -- bash tools/evidence_multi.sh N`); outputs go to gpurun_out/.
set -u
N=${1:-2}
O=gpurun_out
mkdir -p $O
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511"
$TR bench.py --gpus $N --steps 20 --warmup 3 > $O/r2_bench_${N}gpu_weak.json 2> $O/r2_bench_${N}gpu_weak.err
tail -c 400 $O/r2_bench_${N}gpu_weak.json
$TR bench.py --gpus $N --steps 20 --warmup 3 --scaling strong --e2e-steps 4 > $O/r2_bench_${N}gpu_strong.json 2> $O/r2_bench_${N}gpu_strong.err
tail -c 400 $O/r2_bench_${N}gpu_strong.json
# one context, N devices (the product API)
python tools/multidev_e2e.py > $O/r2_multidev_${N}gpu.jsonl 2> $O/r2_multidev_${N}gpu.err; cat $O/r2_multidev_${N}gpu.jsonl
if [ "$N" -le 2 ]; then
  timeout 900 python -m pytest tests/test_gpu_multidevice.py -q 2>&1 | tail -2
  python tools/cli_bench.py --reads 100000 --files 16 --nrec 100000 --S 100 --devices all > $O/r2_cli_bench_${N}gpu.jsonl 2> $O/r2_cli_${N}gpu.err
  cat $O/r2_cli_bench_${N}gpu.jsonl; grep timing $O/r2_cli_${N}gpu.err
fi
