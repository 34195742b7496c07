#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "agc or golden or pruned or overflow" 2>&1 | tail -3
timeout 200 python bench.py --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/c23_dmma.json 2> gpurun_out/c23_dmma.err
timeout 300 python bench.py --steps 3 --warmup 3 --no-cpu-baseline --kpts 8192 --weights damped > gpurun_out/c23_8192.json 2> gpurun_out/c23_8192.err
python - <<'PY'
import json,glob
for f in sorted(glob.glob('gpurun_out/c23_*.json')):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1])
        o=d['roofline_other']
        print(f, round(d['value'],1), 'e2e', round(d['e2e']['value'],1), 'cosine', o['cosine'])
    except Exception as e: print(f, 'ERR', e)
PY
