#!/bin/bash
mkdir -p gpurun_out
for rep in 1 2; do for tpc in 2 3 4; do
GIMS_GEMM_TPC=$tpc timeout 300 python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/c28_tpc${tpc}_$rep.json 2> gpurun_out/c28.err
done; done
python - <<'PY'
import json,glob
for f in sorted(glob.glob('gpurun_out/c28_*.json')):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1])
        print(f, round(d['value'],1), round(d['roofline_other']['gemm']['ms_per_pair'],3))
    except Exception as e: print(f, 'ERR', e)
PY
