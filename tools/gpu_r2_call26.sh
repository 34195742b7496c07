#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_frontend.py -x -q -m gpu -s 2>&1 | tail -12
timeout 600 python bench.py --pipeline --no-cpu-baseline > gpurun_out/c26_pipeline.json 2> gpurun_out/c26_pipeline.err; tail -3 gpurun_out/c26_pipeline.err; cat gpurun_out/c26_pipeline.json
GIMS_HOST_PATCHES=1 timeout 600 python bench.py --pipeline --no-cpu-baseline > gpurun_out/c26_pipeline_host.json 2> gpurun_out/c26_pipeline_host.err; cat gpurun_out/c26_pipeline_host.json
