#!/bin/bash
mkdir -p gpurun_out
N=${1:-1}
if [ "$N" = "1" ]; then
timeout 300 python bench.py --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/c37_sleep_n1.json 2> gpurun_out/c37.err
GIMS_SPIN_SYNC=1 timeout 300 python bench.py --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/c37_spin_n1.json 2>> gpurun_out/c37.err
else
timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29523 bench.py --gpus $N --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/c37_sleep_n$N.json 2> gpurun_out/c37.err
GIMS_SPIN_SYNC=1 timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29523 bench.py --gpus $N --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/c37_spin_n$N.json 2>> gpurun_out/c37.err
fi
python - <<PY
import json,glob
for f in sorted(glob.glob('gpurun_out/c37_*_n$N.json')):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1]); print(f, round(d['value'],1), 'e2e', round(d['e2e']['value'],1), 'single', round(d['e2e']['single_thread_value'],1))
    except Exception as e: print(f,'ERR',e)
PY
