#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py -x -q -m gpu 2>&1 | tail -4
timeout 300 python tools/prof_forward.py 2048 2>&1 | head -3
GIMS_FORK_IMAGES=0 timeout 300 python tools/prof_forward.py 2048 2>&1 | head -3
timeout 200 python bench.py --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/c19_fork.json 2> gpurun_out/c19_fork.err
GIMS_FORK_IMAGES=0 timeout 200 python bench.py --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/c19_nofork.json 2> gpurun_out/c19_nofork.err
python - <<'PY'
import json,glob
for f in sorted(glob.glob('gpurun_out/c19_*.json')):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1])
        print(f, round(d['value'],1), 'e2e', round(d['e2e']['value'],1), 'single-thread', round(d['e2e']['single_thread_value'],1))
    except Exception as e: print(f, 'ERR', e)
PY
