#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "sinkhorn or n4096 or n8192" 2>&1 | tail -4
timeout 120 python tools/sink_trace.py 4096 0 2>&1 | head -6
timeout 120 python tools/sink_trace.py 8192 0 2>&1 | head -6
timeout 120 python tools/sink_trace.py 3000 0 2>&1 | head -6
timeout 300 python bench.py --steps 3 --warmup 3 --no-cpu-baseline --kpts 8192 --weights damped > gpurun_out/c15_8192.json 2> gpurun_out/c15_8192.err
timeout 300 python bench.py --steps 3 --warmup 3 --no-cpu-baseline --kpts 4096 --weights damped > gpurun_out/c15_4096.json 2> gpurun_out/c15_4096.err
python - <<'PY'
import json,glob
for f in sorted(glob.glob('gpurun_out/c15_*.json')):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1])
        o=d['roofline_other']
        print(f, round(d['value'],2), 'e2e', round(d['e2e']['value'],2), 'attn ms', round(d['roofline']['avg_launch_ms'],4), 'gemm', round(o['gemm']['ms_per_pair'],3), 'sink ms', round(o['sinkhorn']['ms_per_pair'],3))
    except Exception as e: print(f, 'ERR', e)
PY
