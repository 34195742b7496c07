#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_frontend.py tests/test_gpu_parity.py -x -q -m gpu -k "frontend or patches or homography or drop_in" 2>&1 | tail -4
timeout 600 python bench.py --pipeline --no-cpu-baseline > gpurun_out/c26_pipeline.json 2> gpurun_out/c26_pipeline.err; tail -3 gpurun_out/c26_pipeline.err; cat gpurun_out/c26_pipeline.json
