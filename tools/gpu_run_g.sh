#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_align_gpu.py -m gpu -x -q 2>&1 | tail -3
timeout 300 python tools/prof_align.py v6c 2>&1 | tail -1
timeout 300 python tools/prof_align.py v6c 2>&1 | tail -1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:align_kernel -c 1 -f -o gpurun_out/r2_align_v6 python tools/prof_align.py --once > gpurun_out/r2g_ncu.log 2>&1
echo "ncu rc=$?"; tail -2 gpurun_out/r2g_ncu.log
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/r2g_pytest.log 2>&1
echo "pytest rc=$?"; tail -4 gpurun_out/r2g_pytest.log
