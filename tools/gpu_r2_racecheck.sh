#!/bin/bash
timeout 1200 compute-sanitizer --tool racecheck --error-exitcode 7 --launch-timeout 900 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "(sinkhorn_vs_oracle and 300-257) or (sinkhorn_vs_oracle and 3100)" > gpurun_out/race_sink.txt 2>&1; echo "racecheck sinkhorn rc=$?"
grep -E "passed|failed|RACECHECK SUMMARY|hazard" gpurun_out/race_sink.txt | sort | uniq -c | head -12
timeout 1200 compute-sanitizer --tool racecheck --error-exitcode 7 --launch-timeout 900 python -m pytest tests/test_gpu_postprocess.py tests/test_gpu_frontend.py -x -q -m gpu -k "300 or 240" > gpurun_out/race_aux.txt 2>&1; echo "racecheck aux rc=$?"
grep -E "passed|failed|RACECHECK SUMMARY|hazard" gpurun_out/race_aux.txt | sort | uniq -c | head -8
