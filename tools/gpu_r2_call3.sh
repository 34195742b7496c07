#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_gemm.py -m gpu -q --tb=short -s -x > gpurun_out/test_gemm.log 2>&1; echo "gemm tests rc=$?"
grep -n "attn layer\|passed\|failed\|FAILED\|Error" gpurun_out/test_gemm.log | tail -40
timeout 1200 python -m pytest tests/test_gpu_parity.py -m gpu -q --tb=short -s > gpurun_out/test_parity.log 2>&1; echo "parity rc=$?"
grep -n "passed\|failed\|FAILED" gpurun_out/test_parity.log | tail; grep -n "^\[fwd\|^\[stages" gpurun_out/test_parity.log | cut -c1-330
for cfg in "2048 random" "4096 damped" "8192 damped"; do
  set -- $cfg
  timeout 600 python bench.py --kpts $1 --weights $2 --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/bench_$1_$2.log 2>&1; echo "bench $1 $2 rc=$?"
  tail -1 gpurun_out/bench_$1_$2.log | cut -c1-200
done
