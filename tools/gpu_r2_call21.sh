#!/bin/bash
timeout 600 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "agc or golden or pruned or overflow" 2>&1 | tail -3
