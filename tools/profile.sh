#!/bin/bash
# ncu evidence for profiles/ (run under gpurun, 1 GPU).  Each command is first
# run plain and must exit 0 before it is repeated under ncu.
#   tools/profile.sh <tag>
set -u
TAG=${1:-r01}
OUT=gpurun_out
mkdir -p $OUT
# (1) the bench line itself, then the launch list of the same command: per-launch
#     device time, cold-cache and serialised -- compare SHARES of the step
python bench.py --steps 3 --warmup 3 > $OUT/${TAG}_bench_n1.json 2> $OUT/${TAG}_bench_n1.err
python bench.py --steps 1 --warmup 3 --no-cpu --no-extras > $OUT/${TAG}_bench_plain.json 2> $OUT/${TAG}_bench_plain.err &&
ncu --metrics gpu__time_duration.sum --clock-control none -s 60 -c 600 --csv \
    --log-file $OUT/${TAG}_launches.csv python bench.py --steps 1 --warmup 3 --no-cpu --no-extras \
    > $OUT/${TAG}_bench_ncu.log 2>&1
# (2) DRAM traffic per launch of the three PCG kernels at the bench workload
#     (single-pass metrics: no replay, no 45 GB save/restore)
python tools/probe.py poisson27 512 0 pcg > $OUT/${TAG}_traffic_plain.log 2>&1 &&
ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum \
    --clock-control none -k regex:'k_spmv_sell|k_pcg_update|k_pcg_pupdate' -s 6 -c 9 --csv \
    --log-file $OUT/${TAG}_traffic_poisson27_512.csv python tools/probe.py poisson27 512 0 pcg \
    > $OUT/${TAG}_traffic_ncu.log 2>&1
# (3) full-set capture of the three PCG kernels (27-point 256^3: same kernels,
#     replay-friendly size)
python tools/probe.py poisson27 256 0 pcg > $OUT/${TAG}_probe_plain.log 2>&1 &&
ncu --set full --clock-control none --import-source on \
    -k regex:'k_spmv_sell|k_pcg_update|k_pcg_pupdate' -s 6 -c 3 \
    -o $OUT/${TAG}_pcg_poisson27_256 -f python tools/probe.py poisson27 256 0 pcg > $OUT/${TAG}_probe_ncu.log 2>&1
ls -la $OUT | tail -12
