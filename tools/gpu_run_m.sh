#!/bin/bash
mkdir -p gpurun_out
LIB=face-recognition-platform_b200/frp_b200/_lib/libfrp_b200.so
cp $LIB /tmp/default.so
cp tools/variants/dual8th8.so $LIB
BISECT_SIGMAS=0.5,1.5,2.6 BISECT_THETAS=0.0,0.3 timeout 300 python tools/align_bisect.py 2>&1 | sed 's/FAIL.*/FAIL/'
cp tools/variants/dual8.so $LIB
BISECT_SIGMAS=2.6,2.8 BISECT_THETAS=0.2,0.3 timeout 300 python tools/align_bisect.py 2>&1 | sed 's/FAIL.*/FAIL/'
timeout 600 python -m pytest tests/test_align_gpu.py -m gpu -x -q 2>&1 | tail -2
for i in 1 2; do timeout 300 python tools/prof_align.py dual8 2>/dev/null | tail -1; done
cp /tmp/default.so $LIB
for i in 1 2; do timeout 300 python tools/prof_align.py default 2>/dev/null | tail -1; done
