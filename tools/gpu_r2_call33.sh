#!/bin/bash
mkdir -p gpurun_out
for s in 12 16 24; do
timeout 300 python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-e2e --streams $s > gpurun_out/c33_s$s.json 2> gpurun_out/c33.err
done
timeout 300 python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-e2e --streams 24 --pairs-per-step 48 > gpurun_out/c33_s24_p48.json 2>> gpurun_out/c33.err
timeout 300 python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-e2e --streams 16 --pairs-per-step 32 > gpurun_out/c33_s16_p32.json 2>> gpurun_out/c33.err
python - <<'PY'
import json,glob
for f in sorted(glob.glob('gpurun_out/c33_*.json')):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1])
        print(f, round(d['value'],1))
    except Exception as e: print(f, 'ERR', e)
PY
