#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "sinkhorn or n4096" 2>&1 | tail -2
timeout 120 python tools/sink_trace.py 4096 0 2>&1 | head -6
timeout 120 python tools/sink_trace.py 3000 0 2>&1 | head -6
