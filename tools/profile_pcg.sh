#!/bin/bash
# full-set ncu capture of the three kernels of one PCG iteration (1 GPU):
#   tools/profile_pcg.sh <tag> <kind> <size>
set -u
TAG=${1:-r01}; KIND=${2:-poisson27}; SIZE=${3:-256}
OUT=gpurun_out; mkdir -p $OUT
python tools/probe.py $KIND $SIZE 0 pcg > $OUT/${TAG}_pcg_${KIND}${SIZE}_plain.log 2>&1 &&
ncu --set full --clock-control none --import-source on \
    -k regex:'k_spmv_sell|k_pcg_update|k_pcg_pupdate' -s 10 -c 6 \
    -o $OUT/${TAG}_pcg_${KIND}${SIZE} -f python tools/probe.py $KIND $SIZE 0 pcg \
    > $OUT/${TAG}_pcg_${KIND}${SIZE}_ncu.log 2>&1
tail -3 $OUT/${TAG}_pcg_${KIND}${SIZE}_ncu.log
