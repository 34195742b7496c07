#!/bin/bash
mkdir -p gpurun_out
timeout 400 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -s -k "bf16" 2>&1 | grep "bf16 variant\|passed\|failed" | tee gpurun_out/c13_bf16.txt
timeout 300 python bench.py --steps 3 --warmup 3 --no-cpu-baseline --kpts 8192 --weights damped --gemm-mode bf16 > gpurun_out/c13_8192_bf16.json 2> gpurun_out/c13_8192_bf16.err
timeout 300 python bench.py --steps 3 --warmup 3 --no-cpu-baseline --kpts 8192 --weights damped > gpurun_out/c13_8192_f16.json 2> gpurun_out/c13_8192_f16.err
timeout 300 python bench.py --steps 3 --warmup 3 --no-cpu-baseline --kpts 4096 --weights damped > gpurun_out/c13_4096_f16.json 2> gpurun_out/c13_4096_f16.err
timeout 300 python bench.py --steps 5 --warmup 3 --no-cpu-baseline --gemm-mode bf16 > gpurun_out/c13_2048_bf16.json 2> gpurun_out/c13_2048_bf16.err
python - <<'PY'
import json,glob
for f in sorted(glob.glob('gpurun_out/c13_*.json')):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1])
        o=d['roofline_other']
        print(f, round(d['value'],2), 'e2e', round(d['e2e']['value'],2), 'attn ms', round(d['roofline']['avg_launch_ms'],4), 'frac', round(d['roofline']['frac'],3), 'gemm', round(o['gemm']['ms_per_pair'],3), 'sink', round(o['sinkhorn']['ms_per_pair'],3))
    except Exception as e: print(f, 'ERR', e)
PY
