#!/bin/bash
timeout 300 python -m pytest tests/test_gpu_postprocess.py -x -q -m gpu -s 2>&1 | tail -15
