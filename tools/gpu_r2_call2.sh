#!/bin/bash
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -q --tb=short -s -x > gpurun_out/test_all.log 2>&1; echo "tests rc=$?"
grep -n "sinkhorn \|passed\|failed\|FAILED" gpurun_out/test_all.log | tail -30
for cfg in "2048 random" "2048 damped" "4096 damped" "8192 damped"; do
  set -- $cfg
  timeout 600 python bench.py --kpts $1 --weights $2 --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/bench_$1_$2.log 2>&1; echo "bench $1 $2 rc=$?"
  tail -1 gpurun_out/bench_$1_$2.log | cut -c1-200
done
for n in 2048 4096 8192; do timeout 120 python tools/sink_trace.py $n 90 > gpurun_out/sink_trace_$n.txt 2>&1; head -16 gpurun_out/sink_trace_$n.txt; done
