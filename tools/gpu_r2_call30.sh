#!/bin/bash
mkdir -p gpurun_out
timeout 120 python tools/sink_trace.py 2048 0 2>&1 | head -9
timeout 600 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "sinkhorn or golden or batched" 2>&1 | tail -2
timeout 300 python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/c30.json 2> gpurun_out/c30.err
python - <<'PY'
import json
d=json.loads(open('gpurun_out/c30.json').read().strip().splitlines()[-1])
print(round(d['value'],1), d['roofline_other']['sinkhorn']['ms_per_pair'])
PY
