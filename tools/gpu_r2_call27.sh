#!/bin/bash
timeout 600 python -m pytest tests/test_gpu_gemm.py -x -q -m gpu 2>&1 | tail -4
GIMS_GEMM_TPC=1 timeout 600 python -m pytest tests/test_gpu_gemm.py -x -q -m gpu -k "f16" 2>&1 | tail -2
GIMS_GEMM_TPC=5 timeout 600 python -m pytest tests/test_gpu_gemm.py -x -q -m gpu -k "f16" 2>&1 | tail -2
