#!/bin/bash
timeout 300 python tools/prof_e2e.py 2>&1 | tail -11
