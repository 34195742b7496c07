#!/bin/bash
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_gemm.py -m gpu -q --tb=short -s > gpurun_out/test_gemm.log 2>&1; echo "gemm tests rc=$?"
grep -n "f16x2 gemm\|passed\|failed\|FAILED\|Error" gpurun_out/test_gemm.log | tail -24
timeout 120 python tools/gemm_trace.py 8192 f16 > gpurun_out/gemm_trace_f16.txt 2>&1; grep -n "us\b\|TFLOP" gpurun_out/gemm_trace_f16.txt | head -8
timeout 120 python tools/gemm_trace.py 8192 tf32 > gpurun_out/gemm_trace_tf32.txt 2>&1; grep -n "us\b\|TFLOP" gpurun_out/gemm_trace_tf32.txt | head -8
