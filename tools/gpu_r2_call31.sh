#!/bin/bash
mkdir -p gpurun_out
for rep in 1 2; do for c in 8 32 64; do
CUDA_DEVICE_MAX_CONNECTIONS=$c timeout 300 python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/c31_conn${c}_$rep.json 2> gpurun_out/c31.err
done; done
CUDA_DEVICE_MAX_CONNECTIONS=32 timeout 300 python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-e2e --kpts 4096 --weights damped > gpurun_out/c31_4096_conn32.json 2>> gpurun_out/c31.err
CUDA_DEVICE_MAX_CONNECTIONS=8 timeout 300 python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-e2e --kpts 4096 --weights damped > gpurun_out/c31_4096_conn8.json 2>> gpurun_out/c31.err
python - <<'PY'
import json,glob
for f in sorted(glob.glob('gpurun_out/c31_*conn*.json')):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1])
        print(f, round(d['value'],1))
    except Exception as e: print(f, 'ERR', e)
PY
