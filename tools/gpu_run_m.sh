#!/bin/bash
mkdir -p gpurun_out
LIB=face-recognition-platform_b200/frp_b200/_lib/libfrp_b200.so
cp $LIB /tmp/default.so
cp tools/variants/dual8th8.so $LIB
BISECT_SIGMAS=0.5,1.0,1.5,2.0,2.6 BISECT_THETAS=0.0,0.3 timeout 600 python tools/align_bisect.py 2>&1 | sed 's/FAIL.*/FAIL/' > gpurun_out/r2m_bisect_th8.log
cat gpurun_out/r2m_bisect_th8.log
cp /tmp/default.so $LIB
