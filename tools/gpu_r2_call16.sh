#!/bin/bash
for i in 1 2 3; do timeout 300 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -s -k "sinkhorn_vs_oracle and 700" 2>&1 | grep "\[sinkhorn\|passed\|failed"; done
git stash -q; python -m gims_b200.build > /dev/null 2>&1; echo "== HEAD build"
for i in 1 2; do timeout 300 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -s -k "sinkhorn_vs_oracle and 700" 2>&1 | grep "\[sinkhorn\|passed\|failed"; done
