#!/bin/bash
mkdir -p gpurun_out
timeout 300 python tools/prof_forward.py 2048 > gpurun_out/c18_prof.txt 2>&1
cat gpurun_out/c18_prof.txt | head -70
