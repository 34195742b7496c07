#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "drop_in or overflow or pruned or batch_of_two or concurrent or homography" 2>&1 | tail -2
for t in 8 12 16; do
timeout 300 python bench.py --steps 5 --warmup 3 --no-cpu-baseline --e2e-threads $t > gpurun_out/c35_t$t.json 2> gpurun_out/c35.err
done
python - <<'PY'
import json,glob
for f in sorted(glob.glob('gpurun_out/c35_*.json')):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1])
        print(f, round(d['value'],1), 'e2e', round(d['e2e']['value'],1), 'single', round(d['e2e']['single_thread_value'],1))
    except Exception as e: print(f, 'ERR', e)
PY
