#!/bin/bash
# Round evidence on ONE B200 (run under gpurun):  tools/evidence.sh <tag>
#  (1) config 2: Nek table incl. the reference's cuSOLVER backend
#  (2) config 5: power-law kernel-selection sweep
#  (3) DRAM traffic of the dominant kernel at the bench workload (single-pass ncu)
set -u
TAG=${1:-r01}
OUT=gpurun_out
mkdir -p $OUT
python tools/nek_table.py 100 5 > $OUT/${TAG}_nek_table.txt 2> $OUT/${TAG}_nek_table.err
tail -9 $OUT/${TAG}_nek_table.txt
timeout 900 python tools/powerlaw_sweep.py 50000000 1 2 3 > $OUT/${TAG}_powerlaw_n1.jsonl 2> $OUT/${TAG}_powerlaw_n1.err
cat $OUT/${TAG}_powerlaw_n1.jsonl | cut -c1-400
python tools/probe.py poisson27 512 0 pcg > $OUT/${TAG}_traffic_plain.log 2>&1 &&
timeout 600 ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum \
    --clock-control none -k regex:'k_spmv_sell|k_pcg_update|k_pcg_pupdate' -s 6 -c 9 --csv \
    --log-file $OUT/${TAG}_traffic_poisson27_512.csv python tools/probe.py poisson27 512 0 pcg \
    > $OUT/${TAG}_traffic_ncu.log 2>&1
tail -12 $OUT/${TAG}_traffic_poisson27_512.csv | cut -c1-300
