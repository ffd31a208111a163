#!/bin/bash
set -u
OUT=gpurun_out
mkdir -p $OUT
timeout 600 python tools/tma_check.py 256 > $OUT/r02d_tma256.log 2>&1; tail -8 $OUT/r02d_tma256.log | cut -c1-600
cp $OUT/tma_check.json $OUT/r02d_tma256.json 2>/dev/null
timeout 900 python -m pytest tests -m gpu -q -k "oracles_solve or full_size_config4 or bench_line" > $OUT/r02d_gputests.log 2>&1; tail -12 $OUT/r02d_gputests.log
timeout 900 python bench.py --steps 1 --warmup 1 > $OUT/r02d_bench.json 2> $OUT/r02d_bench.err; cut -c1-600 $OUT/r02d_bench.json; tail -3 $OUT/r02d_bench.err
