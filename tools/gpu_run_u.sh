#!/bin/bash
# small-gallery kernel: per-warp k-best lists + union ranking for pure top-k
mkdir -p gpurun_out
timeout 200 python tools/prof_c1.py > gpurun_out/r2u_c1.jsonl 2>gpurun_out/r2u_c1.err || echo "c1 failed"
cat gpurun_out/r2u_c1.jsonl
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:match_small -c 30 --csv --log-file gpurun_out/r2u_small_sweep.csv python tools/prof_c1.py --sweep > gpurun_out/r2u_sweep.log 2>&1
grep -v "^==" gpurun_out/r2u_small_sweep.csv | awk -F'","' 'NR>1{print $8, $NF}' | tr -d '"' | paste - - - - 
timeout 900 python -m pytest tests/test_match_gpu.py tests/test_round2_gpu.py -m gpu -x -q -k "small or tiny or empty or topk or face_service or cooldown or cluster or batching" 2>&1 | tail -3
