#!/bin/bash
# ncu evidence for profiles/ (run under gpurun, 1 GPU).  Each command is first
# run plain and must exit 0 before it is repeated under ncu.
#   tools/profile.sh <tag>
set -u
TAG=${1:-r01}
OUT=gpurun_out
mkdir -p $OUT
# (1) launch list of the bench command: per-launch device time, cold-cache and
#     serialised -- compare SHARES of the step, not absolutes
python bench.py --steps 1 --warmup 3 --no-cpu > $OUT/${TAG}_bench_plain.json 2> $OUT/${TAG}_bench_plain.err &&
ncu --metrics gpu__time_duration.sum --clock-control none -s 60 -c 600 --csv \
    --log-file $OUT/${TAG}_launches.csv python bench.py --steps 1 --warmup 3 --no-cpu \
    > $OUT/${TAG}_bench_ncu.log 2>&1
# (2) full-set capture of the three PCG kernels on the SpMV target grid
python tools/probe.py poisson7 256 > $OUT/${TAG}_probe_plain.log 2>&1 &&
ncu --set full --clock-control none --import-source on \
    -k regex:'k_spmv_sell|k_pcg_update|k_pcg_pupdate' -s 30 -c 9 \
    -o $OUT/${TAG}_prof -f python tools/probe.py poisson7 256 > $OUT/${TAG}_probe_ncu.log 2>&1
ls -la $OUT
