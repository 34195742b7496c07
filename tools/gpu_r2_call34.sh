#!/bin/bash
mkdir -p gpurun_out
for st in 5 40 160; do
timeout 600 python bench.py --steps $st --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/c34_steps$st.json 2> gpurun_out/c34.err
done
python - <<'PY'
import json,glob
for f in sorted(glob.glob('gpurun_out/c34_*.json')):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1])
        print(f, round(d['value'],1), d['clocks'])
    except Exception as e: print(f, 'ERR', e)
PY
