#!/bin/bash
# The commands behind profiles/r2_* (run on a B200 box: `gpurun -- bash tools/evidence.sh`); outputs go to gpurun_out/.
# A number printed by a run under ncu is never a bench value: the plain runs come first.
set -u
O=gpurun_out
mkdir -p $O
python bench.py > $O/r2_bench_default.json 2> $O/r2_bench_default.err
tail -c 200 $O/r2_bench_default.json
python bench.py --impl reference --steps 2 --warmup 1 > $O/r2_bench_reference.json 2>/dev/null
# launch list of the same command (shares of the step, cold and serialised)
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $O/r2_launches.csv \
    python bench.py --steps 20 --warmup 3 --no-subrecords --no-cpu-baseline --no-parity --e2e-steps 3 > $O/r2_launches_run.log 2>&1
wc -l $O/r2_launches.csv
# full captures: cfg2 (scan, triage, locate), cfg3 scan, cfg4 filter + scan
python tools/run_once.py --runs 2
timeout 900 ncu --set full --clock-control none --import-source on -k regex:ntl_ -s 3 -c 3 -o $O/r2_kernels_full -f \
    python tools/run_once.py --runs 2 > $O/r2_ncu_a.log 2>&1
python tools/run_once.py --runs 2 --tvr "TTGGG CCAGGG TCAGGG"
timeout 900 ncu --set full --clock-control none -k regex:ntl_scan -s 1 -c 1 -o $O/r2_scan_cfg3 -f \
    python tools/run_once.py --runs 2 --tvr "TTGGG CCAGGG TCAGGG" > $O/r2_ncu_b.log 2>&1
python tools/run_once.py --runs 2 --patterns TTAGGG --no-rc --use-filter --right-edge --telomeric-frac 0.3 --seed 20261022
timeout 900 ncu --set full --clock-control none -k regex:"ntl_scan|ntl_filter" -s 2 -c 2 -o $O/r2_scan_cfg4 -f \
    python tools/run_once.py --runs 2 --patterns TTAGGG --no-rc --use-filter --right-edge --telomeric-frac 0.3 --seed 20261022 > $O/r2_ncu_c.log 2>&1
# host side
PACK_THREADS=1,8,16,20,24,32 python tools/pack_bench.py > $O/r2_pack_bench.txt 2>&1; cat $O/r2_pack_bench.txt
python tools/host_probe.py > $O/r2_host_probe.txt 2>&1; cat $O/r2_host_probe.txt
python tools/cli_bench.py --reads 100000 --files 16 --nrec 10000 100000 --S 100 > $O/r2_cli_bench.jsonl 2> $O/r2_cli.err
cat $O/r2_cli_bench.jsonl; grep timing $O/r2_cli.err
