#!/bin/bash
# The first GPU call of round 2 (one B200, about 3 minutes of box time):
#   gpurun --timeout 400 -- 'bash tools/round2_first_call.sh'
# Everything here was written after round 1's GPU budget was spent; each step is
# independent and writes to gpurun_out/r02a_*.
set -u
OUT=gpurun_out
mkdir -p $OUT
# (1) the software-pipelined SELL kernel: parity on the GPU first, then times beside
#     the plain kernels.  PIPE=1: fp32-stored values only; PIPE=2: fp64 values as well.
B200_SPMV_PIPE=1 timeout 90 python tools/pipe_check.py 192 > $OUT/r02a_pipe1.log 2>&1; tail -2 $OUT/r02a_pipe1.log
cp $OUT/pipe_check.json $OUT/r02a_pipe1.json 2>/dev/null
B200_SPMV_PIPE=2 timeout 90 python tools/pipe_check.py 192 > $OUT/r02a_pipe2.log 2>&1; tail -2 $OUT/r02a_pipe2.log
cp $OUT/pipe_check.json $OUT/r02a_pipe2.json 2>/dev/null
# (2) the plain kernels, same process layout, for the comparison (round 1: f64 0.292 ms,
#     f32 0.306 ms, three-kernel 0.399 ms/it, single-reduction 0.400 ms/it at 192^3)
timeout 60 python tools/variants_probe.py 192 > $OUT/r02a_variants.log 2>&1; tail -1 $OUT/r02a_variants.log | cut -c1-400
# (3) if (1) passed: one full-set capture of the pipelined kernel (27-point 256^3)
if grep -q "parity: ok" $OUT/r02a_pipe1.log; then
  B200_SPMV_PIPE=1 timeout 120 ncu --set full --clock-control none --import-source on \
      -k regex:'k_spmv_sellc32p' -s 4 -c 2 -o $OUT/r02a_sellc32p_poisson27_256 -f \
      python tools/pipe_check.py 256 > $OUT/r02a_pipe_ncu.log 2>&1
fi
ls -la $OUT | grep r02a
# (4) column-blocked SpMV of the power-law operator (config 5): parity, then block widths
timeout 300 python tools/colblock_check.py 50000000 > $OUT/r02a_colblock.log 2>&1; tail -7 $OUT/r02a_colblock.log | cut -c1-300
# (5) the on-chip coarse-grid kernel: plain run, then one full-set capture
timeout 120 python tools/smallprof.py > $OUT/r02a_small_plain.log 2>&1 && \
timeout 300 ncu --set full --clock-control none --import-source on -k regex:'k_pcg_small' -s 1 -c 1 \
    -o $OUT/r02a_pcg_small -f python tools/smallprof.py > $OUT/r02a_small_ncu.log 2>&1
tail -3 $OUT/r02a_small_plain.log
