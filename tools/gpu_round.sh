#!/bin/bash
# One GPU-box visit: parity tests (grouped so a faulting kernel does not poison the rest), smoke, bench, launch list.
# usage: tools/gpu_round.sh [stage ...]   stages: tc agc sink fwd all smoke bench ncu ncufull sanitize
mkdir -p gpurun_out
STAGES="${@:-agc sink fwd smoke bench ncu}"
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/gpu.txt 2>&1
python -c "import os;print('cpus',os.cpu_count())" >> gpurun_out/gpu.txt
grep -m1 "model name" /proc/cpuinfo >> gpurun_out/gpu.txt
for s in $STAGES; do
  case $s in
    tc)    timeout 300 python -m pytest tests/test_gpu_gemm.py -m gpu -q --tb=short -s > gpurun_out/test_tc.log 2>&1; echo "tc rc=$?" ;;
    fwd_simt) GIMS_GEMM_MODE=simt timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -q --tb=short -k "not agc and not sinkhorn" -s > gpurun_out/test_fwd_simt.log 2>&1; echo "fwd_simt rc=$?" ;;
    agc)   timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -q --tb=short -k "agc" -s > gpurun_out/test_agc.log 2>&1; echo "agc rc=$?" ;;
    sink)  timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -q --tb=short -k "sinkhorn" -s > gpurun_out/test_sink.log 2>&1; echo "sink rc=$?" ;;
    fwd)   timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -q --tb=short -k "not agc and not sinkhorn" -s > gpurun_out/test_fwd.log 2>&1; echo "fwd rc=$?" ;;
    all)   timeout 1500 python -m pytest tests -m gpu -q --tb=short > gpurun_out/test_all.log 2>&1; echo "all rc=$?" ;;
    smoke) timeout 300 python __graft_entry__.py smoke > gpurun_out/smoke.log 2>&1; echo "smoke rc=$?" ;;
    bench) timeout 900 python bench.py --steps 3 --warmup 3 > gpurun_out/bench.log 2>&1; echo "bench rc=$?"; tail -1 gpurun_out/bench.log ;;
    ncu)   timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 1200 --csv --log-file gpurun_out/launches.csv \
             python bench.py --steps 1 --warmup 3 --pairs-per-step 1 --streams 1 --pool 2 --no-cpu-baseline --no-e2e > gpurun_out/ncu_bench.log 2>&1; echo "ncu rc=$?" ;;
    ncufull) B="python bench.py --steps 1 --warmup 3 --pairs-per-step 1 --streams 1 --pool 2 --no-cpu-baseline --no-e2e"
           timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_attention_tc -s 40 -c 1 -f -o gpurun_out/ncu_attention $B > gpurun_out/ncufull_attn.log 2>&1; echo "ncufull attn rc=$?"
           timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_gemm_tc -s 158 -c 4 -f -o gpurun_out/ncu_gemm_layer $B > gpurun_out/ncufull_gemm.log 2>&1; echo "ncufull gemm rc=$?"
           timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_sinkhorn -s 2 -c 1 -f -o gpurun_out/ncu_sinkhorn $B > gpurun_out/ncufull_sink.log 2>&1; echo "ncufull sink rc=$?" ;;
    sanitize) timeout 900 compute-sanitizer --tool memcheck --error-exitcode 9 python __graft_entry__.py smoke > gpurun_out/sanitize.log 2>&1; echo "sanitize rc=$?" ;;
  esac
done
tail -5 gpurun_out/test_*.log 2>/dev/null | tail -40
