#!/bin/bash
mkdir -p gpurun_out
GIMS_GEMM_TPC=4 timeout 120 python tools/gemm_trace.py 8192 f16 2>&1 | head -6 | tee gpurun_out/c12_trace_tpc4.txt
GIMS_GEMM_TPC=2 timeout 120 python tools/gemm_trace.py 8192 f16 2>&1 | grep us/launch
timeout 300 python -m pytest tests/test_gpu_gemm.py -x -q -m gpu 2>&1 | tail -3
for i in 1 2; do
timeout 200 python bench.py --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/c12_persist$i.json 2> gpurun_out/c12_persist.err
GIMS_GEMM_PERSIST=0 timeout 200 python bench.py --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/c12_old$i.json 2> gpurun_out/c12_old.err
done
python - <<'PY'
import json,glob
for f in sorted(glob.glob('gpurun_out/c12_*.json')):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1])
        o=d['roofline_other']
        print(f, round(d['value'],1), 'e2e', round(d['e2e']['value'],1), 'single-thread', round(d['e2e']['single_thread_value'],1), 'attn ms', round(d['roofline']['avg_launch_ms'],4), 'gemm', round(o['gemm']['ms_per_pair'],3), 'sink', round(o['sinkhorn']['ms_per_pair'],3))
    except Exception as e: print(f, 'ERR', e)
PY
