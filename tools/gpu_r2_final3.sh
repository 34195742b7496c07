#!/bin/bash
mkdir -p gpurun_out
timeout 1800 python -m pytest tests -x -q -m gpu 2>&1 | tail -3
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
timeout 600 python bench.py > gpurun_out/final3_default.json 2> gpurun_out/final3_default.err; echo "bench rc=$?"
python - <<'PY'
import json
d=json.loads(open('gpurun_out/final3_default.json').read().strip().splitlines()[-1])
print(round(d['value'],1), 'e2e', round(d['e2e']['value'],1), 'roofline', round(d['roofline']['frac'],3), 'cpu', d['cpu_baseline']['value'], 'launches', d['gpu_launches'])
PY
