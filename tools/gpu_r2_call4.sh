#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -q --tb=short -s -k "bf16 or homography or overflow or pruned or batch_of" > gpurun_out/test_new.log 2>&1; echo "new tests rc=$?"
grep -n "bf16 variant\|passed\|failed\|FAILED\|Error" gpurun_out/test_new.log | tail -12
timeout 300 python bench.py --pipeline --kpts 2048 --weights damped --steps 4 > gpurun_out/bench_pipeline.log 2>&1; echo "pipeline rc=$?"; tail -1 gpurun_out/bench_pipeline.log
timeout 300 python bench.py --gemm-mode bf16 --steps 3 --warmup 3 --no-cpu-baseline --weights damped > gpurun_out/bench_2048_bf16.log 2>&1; echo "bf16 rc=$?"; tail -1 gpurun_out/bench_2048_bf16.log | cut -c1-160
timeout 300 python bench.py --gemm-mode tf32 --steps 3 --warmup 3 --no-cpu-baseline --weights damped > gpurun_out/bench_2048_tf32.log 2>&1; echo "tf32 rc=$?"; tail -1 gpurun_out/bench_2048_tf32.log | cut -c1-160
