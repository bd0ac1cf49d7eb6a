#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_round2_gpu.py -m gpu -x -q -k "pure_topk" > gpurun_out/r2i_pytest.log 2>&1
grep -n "Error\|assert\|^E" gpurun_out/r2i_pytest.log | head -30
