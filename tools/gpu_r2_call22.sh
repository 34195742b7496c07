#!/bin/bash
mkdir -p gpurun_out
B="python bench.py --steps 1 --warmup 3 --pairs-per-step 2 --streams 1 --pool 2 --pairs-per-launch 2 --no-cpu-baseline --no-e2e"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_cosine_fp64 -s 6 -c 1 -f -o gpurun_out/ncu_cosine $B > gpurun_out/ncufull_cos.log 2>&1; echo "rc=$?"
ncu -i gpurun_out/ncu_cosine.ncu-rep --page details 2>/dev/null | grep -i -E "duration|fp64|pipe|issue|stall|eligible|warp cycles|registers|theoretical occ|achieved occ|No Eligible|One or More" | head -50
