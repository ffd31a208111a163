#!/bin/bash
set -u
OUT=gpurun_out
mkdir -p $OUT
timeout 300 tools/micro/bulk_probe > $OUT/r02c_bulk_probe.txt 2>&1; cat $OUT/r02c_bulk_probe.txt
timeout 1200 python -m pytest tests -m gpu -q > $OUT/r02c_gputests.log 2>&1; tail -15 $OUT/r02c_gputests.log
