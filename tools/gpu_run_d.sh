#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_align_gpu.py -m gpu -x -q > gpurun_out/r2d_pytest.log 2>&1
echo "pytest rc=$?"; tail -5 gpurun_out/r2d_pytest.log
CUDA_LAUNCH_BLOCKING=1 timeout 600 python tools/align_bisect.py > gpurun_out/r2d_bisect.log 2>&1
cat gpurun_out/r2d_bisect.log
timeout 300 python tools/prof_align.py v6b > gpurun_out/r2d_prof.json 2> gpurun_out/r2d_prof.err
cat gpurun_out/r2d_prof.json; tail -2 gpurun_out/r2d_prof.err
