#!/bin/bash
mkdir -p gpurun_out
timeout 300 python bench.py --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/c24_lanes2.json 2> gpurun_out/c24_lanes2.err
GIMS_COOP_LANES=1 timeout 300 python bench.py --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/c24_lanes1.json 2> gpurun_out/c24_lanes1.err
python - <<'PY'
import json,glob
for f in sorted(glob.glob('gpurun_out/c24_lanes*.json')):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1])
        print(f, round(d['value'],1), 'e2e', round(d['e2e']['value'],1))
    except Exception as e: print(f, 'ERR', e)
PY
