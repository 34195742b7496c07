#!/bin/bash
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_gemm.py -x -q -m gpu 2>&1 | tail -8 > gpurun_out/c11_gemm_tests.txt
tail -4 gpurun_out/c11_gemm_tests.txt
for tpc in 1 2 4; do
  echo "== persistent TPC=$tpc rows 8192"; GIMS_GEMM_TPC=$tpc timeout 120 python tools/gemm_trace.py 8192 f16 2>&1 | grep "us/launch"
done
echo "== one tile per CTA (old kernel) rows 8192"; GIMS_GEMM_PERSIST=0 timeout 120 python tools/gemm_trace.py 8192 f16 2>&1 | grep "us/launch"
echo "== persistent rows 4096"; timeout 120 python tools/gemm_trace.py 4096 f16 2>&1 | grep "us/launch"
echo "== old rows 4096"; GIMS_GEMM_PERSIST=0 timeout 120 python tools/gemm_trace.py 4096 f16 2>&1 | grep "us/launch"
timeout 400 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "golden or batched" 2>&1 | tail -5
timeout 200 python bench.py --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/c11_persist.json 2> gpurun_out/c11_persist.err
GIMS_GEMM_PERSIST=0 timeout 200 python bench.py --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/c11_old.json 2> gpurun_out/c11_old.err
python - <<'PY'
import json,glob
for f in sorted(glob.glob('gpurun_out/c11_*.json')):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1])
        o=d['roofline_other']
        print(f, round(d['value'],1), 'e2e', round(d['e2e']['value'],1), 'single-thread', round(d['e2e']['single_thread_value'],1), 'attn ms', round(d['roofline']['avg_launch_ms'],4), 'gemm', round(o['gemm']['ms_per_pair'],3), 'sink', round(o['sinkhorn']['ms_per_pair'],3))
    except Exception as e: print(f, 'ERR', e)
PY
