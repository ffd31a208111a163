#!/bin/bash
# bash tools/pl_multi.sh N   (under gpurun --gpus N): column-blocked power-law SpMV on N ranks
N=${1:-2}
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
timeout 600 $TR --master-port 29511 tests/dist_check.py > gpurun_out/pl_dist$N.log 2>&1; tail -5 gpurun_out/pl_dist$N.log | cut -c1-300
POWERLAW_SWEEP_ONLY=auto,auto_colblock timeout 600 $TR --master-port 29512 tools/powerlaw_sweep.py 50000000 1 > gpurun_out/pl_sweep$N.log 2>&1
python - <<PY
import json
for l in open('gpurun_out/pl_sweep$N.log'):
    if l.startswith('{'):
        d = json.loads(l)
        print(d.get('selection'), d.get('n_gpus'), d.get('ms_per_spmv'), d.get('algorithmic_gbs'), d.get('error'),
              d.get('rank0', {}).get('col_blocks'), d.get('rank0', {}).get('n_halo'))
PY
tail -3 gpurun_out/pl_sweep$N.log | cut -c1-200
