#!/bin/bash
# one-GPU evidence call: bash tools/gpu_n1.sh TAG
set -u
TAG=${1:-r02n}
OUT=gpurun_out; mkdir -p $OUT
timeout 1500 python bench.py --steps 3 --warmup 3 > $OUT/${TAG}_bench_n1.json 2> $OUT/${TAG}_bench_n1.err; cut -c1-300 $OUT/${TAG}_bench_n1.json; tail -3 $OUT/${TAG}_bench_n1.err
OMP_NUM_THREADS=1 timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > $OUT/${TAG}_ref_n1.json 2> $OUT/${TAG}_ref_n1.err; cut -c1-200 $OUT/${TAG}_ref_n1.json
timeout 1200 python -m pytest tests -m gpu -q > $OUT/${TAG}_gputests.log 2>&1; tail -4 $OUT/${TAG}_gputests.log
