#!/bin/bash
mkdir -p gpurun_out
timeout 120 python tools/sink_trace.py 2048 0 > gpurun_out/c10_trace_half.txt 2>&1
GIMS_SINKHORN_GRID=full timeout 120 python tools/sink_trace.py 2048 0 > gpurun_out/c10_trace_full.txt 2>&1
timeout 400 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "sinkhorn or batched or golden or concurrent" 2>&1 | tail -15 > gpurun_out/c10_tests.txt
timeout 200 python bench.py --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/c10_half.json 2> gpurun_out/c10_half.err
GIMS_SINKHORN_GRID=full timeout 200 python bench.py --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/c10_full.json 2> gpurun_out/c10_full.err
head -16 gpurun_out/c10_trace_half.txt; head -8 gpurun_out/c10_trace_full.txt; tail -4 gpurun_out/c10_tests.txt
python - <<'PY'
import json,glob
for f in sorted(glob.glob('gpurun_out/c10_*.json')):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1])
        o=d['roofline_other']
        print(f, round(d['value'],1), 'e2e', round(d['e2e']['value'],1), 'single-thread', round(d['e2e']['single_thread_value'],1), 'attn ms', round(d['roofline']['avg_launch_ms'],4), 'gemm', round(o['gemm']['ms_per_pair'],3), 'sink', round(o['sinkhorn']['ms_per_pair'],3))
    except Exception as e: print(f, 'ERR', e)
PY
