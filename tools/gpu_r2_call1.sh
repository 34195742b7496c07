#!/bin/bash
# round 2, call 1: everything that has never run (4096 / 8192 goldens, streamed Sinkhorn, both weight sets)
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/gpu.txt 2>&1
python -c "import os;print('cpus',os.cpu_count())" >> gpurun_out/gpu.txt
grep -m1 "model name" /proc/cpuinfo >> gpurun_out/gpu.txt
timeout 1200 python -m pytest tests -m gpu -q --tb=short -s > gpurun_out/test_all.log 2>&1; echo "tests rc=$?"
tail -3 gpurun_out/test_all.log
for cfg in "2048 random" "2048 damped" "4096 random" "4096 damped" "8192 random" "8192 damped"; do
  set -- $cfg
  timeout 600 python bench.py --kpts $1 --weights $2 --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/bench_$1_$2.log 2>&1; echo "bench $1 $2 rc=$?"
  tail -1 gpurun_out/bench_$1_$2.log | cut -c1-400
done
for k in 2048 4096 8192; do
  timeout 400 python bench.py --impl reference --kpts $k --steps 2 --warmup 1 > gpurun_out/ref_$k.log 2>&1; echo "ref $k rc=$?"
  tail -1 gpurun_out/ref_$k.log | cut -c1-200
done
