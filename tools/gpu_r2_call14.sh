#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "sinkhorn or n4096 or n8192" 2>&1 | tail -6 | tee gpurun_out/c14_tests.txt
timeout 120 python tools/sink_trace.py 8192 0 2>&1 | head -8
timeout 120 python tools/sink_trace.py 4096 0 2>&1 | head -8
timeout 300 python bench.py --steps 3 --warmup 3 --no-cpu-baseline --kpts 8192 --weights damped > gpurun_out/c14_8192.json 2> gpurun_out/c14_8192.err
timeout 300 python bench.py --steps 3 --warmup 3 --no-cpu-baseline --kpts 4096 --weights damped > gpurun_out/c14_4096.json 2> gpurun_out/c14_4096.err
python - <<'PY'
import json,glob
for f in sorted(glob.glob('gpurun_out/c14_*.json')):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1])
        o=d['roofline_other']
        print(f, round(d['value'],2), 'e2e', round(d['e2e']['value'],2), 'attn ms', round(d['roofline']['avg_launch_ms'],4), 'frac', round(d['roofline']['frac'],3), 'gemm', round(o['gemm']['ms_per_pair'],3), 'sink', o['sinkhorn'])
    except Exception as e: print(f, 'ERR', e)
PY
