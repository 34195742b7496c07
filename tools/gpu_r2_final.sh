#!/bin/bash
# round 2, final measurements on one B200: every BASELINE config + the launch list for profiles/
mkdir -p gpurun_out
run() { name=$1; shift; timeout 600 python bench.py "$@" > gpurun_out/final_$name.json 2> gpurun_out/final_$name.err; echo "$name rc=$?"; }
run default
run damped --weights damped --no-cpu-baseline
run bf16 --gemm-mode bf16 --no-cpu-baseline
run tf32 --gemm-mode tf32 --no-cpu-baseline
run 4096 --kpts 4096 --weights damped --steps 3 --no-cpu-baseline
run 8192 --kpts 8192 --weights damped --steps 3 --no-cpu-baseline
run 8192_bf16 --kpts 8192 --weights damped --steps 3 --no-cpu-baseline --gemm-mode bf16
run pipeline --pipeline --no-cpu-baseline
run reference --impl reference --steps 2 --warmup 1
B="python bench.py --steps 1 --warmup 3 --pairs-per-step 2 --streams 1 --pool 2 --pairs-per-launch 2 --no-cpu-baseline --no-e2e"
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 1600 --csv --log-file gpurun_out/launches.csv $B > gpurun_out/ncu_bench.log 2>&1; echo "ncu list rc=$?"
timeout 120 python tools/sink_trace.py 2048 0 > gpurun_out/final_sink_trace_2048.txt 2>&1
timeout 120 python tools/sink_trace.py 4096 0 > gpurun_out/final_sink_trace_4096.txt 2>&1
timeout 120 python tools/sink_trace.py 8192 0 > gpurun_out/final_sink_trace_8192.txt 2>&1
GIMS_GEMM_TPC=4 timeout 120 python tools/gemm_trace.py 8192 f16 > gpurun_out/final_gemm_trace_tpc4.txt 2>&1
timeout 120 python tools/gemm_trace.py 8192 f16 > gpurun_out/final_gemm_trace.txt 2>&1
python - <<'PY'
import json,glob
for f in sorted(glob.glob('gpurun_out/final_*.json')):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1])
        print(f, d.get('metric'), round(d['value'],3), 'e2e', round(d['e2e']['value'],3), 'roof', d.get('roofline',{}).get('frac'))
    except Exception as e: print(f, 'ERR', e)
PY
