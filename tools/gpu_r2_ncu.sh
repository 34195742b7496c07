#!/bin/bash
mkdir -p gpurun_out
B="python bench.py --steps 1 --warmup 3 --pairs-per-step 2 --streams 1 --pool 2 --pairs-per-launch 2 --no-cpu-baseline --no-e2e"
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 1600 --csv --log-file gpurun_out/launches.csv $B > gpurun_out/ncu_bench.log 2>&1; echo "ncu list rc=$?"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_attention_f16 -s 40 -c 1 -f -o gpurun_out/ncu_attention $B > gpurun_out/ncufull_attn.log 2>&1; echo "ncufull attn rc=$?"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_gemm_tc -s 158 -c 4 -f -o gpurun_out/ncu_gemm_layer $B > gpurun_out/ncufull_gemm.log 2>&1; echo "ncufull gemm rc=$?"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_sinkhorn_reg -s 2 -c 1 -f -o gpurun_out/ncu_sinkhorn $B > gpurun_out/ncufull_sink.log 2>&1; echo "ncufull sink rc=$?"
B8="python bench.py --kpts 8192 --weights damped --steps 1 --warmup 3 --pairs-per-step 2 --streams 1 --pool 2 --pairs-per-launch 1 --no-cpu-baseline --no-e2e"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_sinkhorn_stream -s 2 -c 1 -f -o gpurun_out/ncu_sinkhorn_stream8192 $B8 > gpurun_out/ncufull_sink8192.log 2>&1; echo "ncufull sink stream rc=$?"
ls -la gpurun_out/*.ncu-rep
