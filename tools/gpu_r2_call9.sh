#!/bin/bash
mkdir -p gpurun_out
for ppl in 1 2 4; do
  timeout 200 python bench.py --steps 5 --warmup 3 --no-cpu-baseline --pairs-per-launch $ppl > gpurun_out/c9_ppl$ppl.json 2> gpurun_out/c9_ppl$ppl.err
done
GIMS_SINKHORN_DUAL=0 timeout 200 python bench.py --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/c9_single.json 2> gpurun_out/c9_single.err
GIMS_SINKHORN_DUAL=0 timeout 200 python bench.py --steps 5 --warmup 3 --no-cpu-baseline --pairs-per-launch 4 > gpurun_out/c9_single4.json 2> gpurun_out/c9_single4.err
python - <<'PY'
import json,glob
for f in sorted(glob.glob('gpurun_out/c9_*.json')):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1])
        o=d['roofline_other']
        print(f, round(d['value'],1), 'e2e', round(d['e2e']['value'],1), 'attn ms', round(d['roofline']['avg_launch_ms'],4), 'gemm', round(o['gemm']['ms_per_pair'],3), 'sink', round(o['sinkhorn']['ms_per_pair'],3))
    except Exception as e: print(f, 'ERR', e)
PY
