#!/bin/bash
timeout 900 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "agc or golden or pruned or overflow or batch" 2>&1 | tail -2
timeout 300 python tools/prof_forward.py 2048 2>&1 | head -2
timeout 300 python bench.py --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/cc.json 2> gpurun_out/cc.err
python - <<'PY'
import json
d=json.loads(open('gpurun_out/cc.json').read().strip().splitlines()[-1]); print(round(d['value'],1), 'e2e', round(d['e2e']['value'],1), 'single', round(d['e2e']['single_thread_value'],1))
PY
B="python bench.py --steps 1 --warmup 3 --pairs-per-step 2 --streams 1 --pool 2 --pairs-per-launch 2 --no-cpu-baseline --no-e2e"
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 1700 --csv --log-file gpurun_out/launches.csv $B > gpurun_out/ncu_bench.log 2>&1
python tools/launch_summary.py gpurun_out/launches.csv 2>/dev/null | grep -E "launches|k_cc|k_cosine|k_sel_hist"
