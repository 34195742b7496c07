#!/bin/bash
echo "== blocking event"; timeout 300 python tools/prof_e2e.py 2>&1 | tail -7
echo "== spin"; GIMS_SPIN_SYNC=1 timeout 300 python tools/prof_e2e.py 2>&1 | tail -4
timeout 300 python bench.py --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/c25_block.json 2> gpurun_out/c25_block.err
GIMS_SPIN_SYNC=1 timeout 300 python bench.py --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/c25_spin.json 2> gpurun_out/c25_spin.err
timeout 300 python bench.py --steps 5 --warmup 3 --no-cpu-baseline --e2e-threads 12 > gpurun_out/c25_block12.json 2> gpurun_out/c25_block12.err
python - <<'PY'
import json,glob
for f in sorted(glob.glob('gpurun_out/c25_*.json')):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1])
        print(f, round(d['value'],1), 'e2e', round(d['e2e']['value'],1), 'single', round(d['e2e']['single_thread_value'],1))
    except Exception as e: print(f, 'ERR', e)
PY
