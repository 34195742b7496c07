#!/bin/bash
mkdir -p gpurun_out
echo "== two chains"; timeout 300 python -m pytest tests/test_gpu_gemm.py -q -m gpu -s -k "f16_gemm" 2>&1 | grep "f16x2 gemm" | awk '{print $NF, $4,$5,$6,$7,$8}' | head -12
echo "== single chain"; GIMS_GEMM_SINGLE_CHAIN=1 timeout 300 python -m pytest tests/test_gpu_gemm.py -q -m gpu -s -k "f16_gemm" 2>&1 | grep "f16x2 gemm\|passed\|failed" | awk '{print $NF, $4,$5,$6,$7,$8}' | head -14
GIMS_GEMM_SINGLE_CHAIN=1 timeout 600 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -s -k "golden or attention" 2>&1 | grep -E "scores|passed|failed" | head -12
timeout 600 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -s -k "golden" 2>&1 | grep -E "scores" | head -10
GIMS_GEMM_SINGLE_CHAIN=1 timeout 300 python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/c36_single.json 2> gpurun_out/c36.err
timeout 300 python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/c36_two.json 2>> gpurun_out/c36.err
python - <<'PY'
import json,glob
for f in sorted(glob.glob('gpurun_out/c36_*.json')):
    d=json.loads(open(f).read().strip().splitlines()[-1]); print(f, round(d['value'],1), round(d['roofline_other']['gemm']['ms_per_pair'],3))
PY
