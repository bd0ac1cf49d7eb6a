#!/bin/bash
mkdir -p gpurun_out
LIB=face-recognition-platform_b200/frp_b200/_lib/libfrp_b200.so
for st in "2.8 0.25" "1.9 0.25"; do
  timeout 120 python tools/align_bisect.py one $st 2>&1 | grep -v "^$" | tail -3
done
timeout 600 python -m pytest tests/test_align_gpu.py -m gpu -x -q 2>&1 | tail -4
timeout 300 python tools/prof_align.py v6slack 2>&1 | tail -1
timeout 300 python tools/prof_align.py v6slack 2>&1 | tail -1
