#!/bin/bash
# final verification on one B200: the whole GPU suite, smoke, the default bench line and the reference arm
mkdir -p gpurun_out
timeout 1800 python -m pytest tests -x -q -m gpu 2>&1 | tail -3 | tee gpurun_out/final2_tests.txt
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
timeout 600 python bench.py > gpurun_out/final2_default.json 2> gpurun_out/final2_default.err; echo "bench rc=$?"
timeout 600 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/final2_reference.json 2> gpurun_out/final2_reference.err; echo "ref rc=$?"
timeout 600 python bench.py --kpts 4096 --weights damped --steps 3 --no-cpu-baseline > gpurun_out/final2_4096.json 2> gpurun_out/final2_4096.err
timeout 600 python bench.py --kpts 8192 --weights damped --steps 3 --no-cpu-baseline > gpurun_out/final2_8192.json 2> gpurun_out/final2_8192.err
timeout 600 python bench.py --gemm-mode bf16 --no-cpu-baseline > gpurun_out/final2_bf16.json 2> gpurun_out/final2_bf16.err
timeout 600 python bench.py --pipeline --no-cpu-baseline > gpurun_out/final2_pipeline.json 2> gpurun_out/final2_pipeline.err
python - <<'PY'
import json,glob
for f in sorted(glob.glob('gpurun_out/final2_*.json')):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1])
        print(f, d.get('metric'), round(d['value'],3), 'e2e', (d.get('e2e') or {}).get('value'), 'roof', (d.get('roofline') or {}).get('frac'))
    except Exception as e: print(f, 'ERR', e)
PY
