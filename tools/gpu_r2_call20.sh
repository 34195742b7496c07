#!/bin/bash
timeout 120 python tools/sink_trace.py 2048 0 2>&1 | head -9
timeout 120 python tools/sink_trace.py 4096 0 2>&1 | head -5
