#!/bin/bash
# compute-sanitizer memcheck: smoke, the auxiliary rows, then the streamed Sinkhorn / persistent GEMM / batched forward
mkdir -p gpurun_out
timeout 900 compute-sanitizer --tool memcheck --error-exitcode 7 --launch-timeout 600 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" > gpurun_out/sanitize_smoke.txt 2>&1; echo "memcheck smoke rc=$?"
timeout 1200 compute-sanitizer --tool memcheck --error-exitcode 7 --launch-timeout 600 python -m pytest tests/test_gpu_postprocess.py tests/test_gpu_frontend.py -x -q -m gpu -k "300 or 1-5 or 240" > gpurun_out/sanitize_aux.txt 2>&1; echo "memcheck aux rc=$?"
timeout 2400 compute-sanitizer --tool memcheck --error-exitcode 7 --launch-timeout 900 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "n4096_damped or batched_pairs or (sinkhorn_vs_oracle and 3100) or (sinkhorn_vs_oracle and 300-257) or all_pruned or overflow" > gpurun_out/sanitize_big.txt 2>&1; echo "memcheck big rc=$?"
tail -6 gpurun_out/sanitize_big.txt
timeout 1200 compute-sanitizer --tool memcheck --error-exitcode 7 --launch-timeout 900 python -m pytest tests/test_gpu_gemm.py -x -q -m gpu -k "f16_gemm" > gpurun_out/sanitize_gemm.txt 2>&1; echo "memcheck gemm rc=$?"
tail -4 gpurun_out/sanitize_gemm.txt
