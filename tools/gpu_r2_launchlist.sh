#!/bin/bash
mkdir -p gpurun_out
B="python bench.py --steps 1 --warmup 3 --pairs-per-step 2 --streams 1 --pool 2 --pairs-per-launch 2 --no-cpu-baseline --no-e2e"
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 1700 --csv --log-file gpurun_out/launches.csv $B > gpurun_out/ncu_bench.log 2>&1; echo "ncu list rc=$?"
python tools/launch_summary.py gpurun_out/launches.csv | head -12
