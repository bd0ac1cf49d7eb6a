#!/bin/bash
# small-gallery kernel: query in the kernel parameters + one-warp ranking of few candidates
mkdir -p gpurun_out
: > gpurun_out/r2t_c1.jsonl
for i in 1 2; do
  timeout 200 python tools/prof_c1.py >> gpurun_out/r2t_c1.jsonl 2>gpurun_out/r2t_c1.err || echo "c1 failed"
done
cat gpurun_out/r2t_c1.jsonl
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:match_small -c 30 --csv --log-file gpurun_out/r2t_small_sweep.csv python tools/prof_c1.py --sweep > gpurun_out/r2t_sweep.log 2>&1
grep -v "^==" gpurun_out/r2t_small_sweep.csv | awk -F'","' 'NR>1{print $8, $NF}' | tr -d '"' | paste - - - - 
timeout 900 python -m pytest tests/test_match_gpu.py tests/test_round2_gpu.py -m gpu -x -q -k "not two_devices" 2>&1 | tail -3
