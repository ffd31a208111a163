#!/bin/bash
# multi-GPU call: bash tools/gpu_multi.sh N TAG   (under gpurun --gpus N)
set -u
N=${1:-2}; TAG=${2:-r02m}
OUT=gpurun_out; mkdir -p $OUT
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
timeout 900 $TR --master-port 29511 tests/dist_check.py > $OUT/${TAG}_dist${N}.log 2>&1; tail -4 $OUT/${TAG}_dist${N}.log | cut -c1-600
timeout 900 $TR --master-port 29512 bench.py --gpus $N --steps 3 --warmup 2 > $OUT/${TAG}_bench_n${N}.json 2> $OUT/${TAG}_bench_n${N}.err; cut -c1-400 $OUT/${TAG}_bench_n${N}.json; tail -2 $OUT/${TAG}_bench_n${N}.err
B200_HALO=nccl timeout 900 $TR --master-port 29513 bench.py --gpus $N --steps 3 --warmup 2 --no-extras > $OUT/${TAG}_bench_n${N}_halo_nccl.json 2> $OUT/${TAG}_bench_n${N}_halo_nccl.err; cut -c1-200 $OUT/${TAG}_bench_n${N}_halo_nccl.json
B200_ALLREDUCE=nccl timeout 900 $TR --master-port 29516 bench.py --gpus $N --steps 3 --warmup 2 --no-extras > $OUT/${TAG}_bench_n${N}_all_nccl.json 2> $OUT/${TAG}_bench_n${N}_all_nccl.err; cut -c1-200 $OUT/${TAG}_bench_n${N}_all_nccl.json
timeout 600 python -m pytest tests/test_gpu_host.py -m gpu -q -k synthetic_and_multi_gpu > $OUT/${TAG}_hostthreads_n${N}.log 2>&1; tail -2 $OUT/${TAG}_hostthreads_n${N}.log
OMP_NUM_THREADS=1 timeout 600 $TR --master-port 29514 bench.py --impl reference --gpus $N --steps 3 --warmup 1 > $OUT/${TAG}_ref_n${N}.json 2> $OUT/${TAG}_ref_n${N}.err; cut -c1-300 $OUT/${TAG}_ref_n${N}.json
