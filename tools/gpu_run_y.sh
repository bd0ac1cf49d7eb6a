#!/bin/bash
# host-call glue: one result buffer + buffer-protocol address look-ups
mkdir -p gpurun_out
timeout 100 python tools/prof_c1.py > gpurun_out/r2y_c1.jsonl 2>gpurun_out/r2y_c1.err || echo "c1 failed"
cat gpurun_out/r2y_c1.jsonl
timeout 200 python -m pytest tests/test_match_gpu.py tests/test_round2_gpu.py -m gpu -x -q -k "small or face_service or empty or tie or lattice or host" 2>&1 | tail -2
