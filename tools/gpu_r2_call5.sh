#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -q --tb=short -x -k "batched or concurrent or golden" > gpurun_out/test_batch.log 2>&1; echo "batch tests rc=$?"
grep -n "passed\|failed\|FAILED\|Error\|assert" gpurun_out/test_batch.log | tail -12
for ppl in 1 2 3 4; do
  timeout 300 python bench.py --pairs-per-launch $ppl --steps 3 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/bench_ppl$ppl.log 2>&1; echo "ppl $ppl rc=$?"
  python - <<PY
import json
d=json.loads(open('gpurun_out/bench_ppl$ppl.log').read().strip().split('\n')[-1])
r=d['roofline']; o=d['roofline_other']
print('ppl $ppl value %.1f | attn %.1f us/launch %.1f TF frac %.3f | gemm %.2f ms/pair %.1f TF | sink %.3f' % (d['value'], r['avg_launch_ms']*1e3, r['achieved'], r['frac'], o['gemm']['ms_per_pair'], o['gemm']['achieved'], o['sinkhorn']['ms_per_pair']))
PY
done
