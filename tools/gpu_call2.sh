#!/bin/bash
# round 2, second GPU call: the bulk-copy-fed SpMV kernel, the GPU test suite with the
# 1e-10 bars, a short bench line
set -u
OUT=gpurun_out
mkdir -p $OUT
timeout 600 python tools/tma_check.py 256 > $OUT/r02b_tma256.log 2>&1; tail -9 $OUT/r02b_tma256.log | cut -c1-700
cp $OUT/tma_check.json $OUT/r02b_tma256.json 2>/dev/null
timeout 600 python -m pytest tests -m gpu -x -q > $OUT/r02b_gputests.log 2>&1; tail -5 $OUT/r02b_gputests.log
timeout 600 python bench.py --steps 1 --warmup 1 --no-cpu > $OUT/r02b_bench.json 2> $OUT/r02b_bench.err; cut -c1-1500 $OUT/r02b_bench.json
