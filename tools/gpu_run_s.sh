#!/bin/bash
# L2 epilogue with packed subtraction (query operand 2q) + small-gallery kernel (cluster, 512 threads, one staging read)
mkdir -p gpurun_out
: > gpurun_out/r2s_l2.txt
for i in 1 2; do
  timeout 300 python tools/prof_match.py --rows 4000000 --dim 128 --metric l2 --batch 256 --iters 6 >> gpurun_out/r2s_l2.txt 2>&1
  timeout 300 python tools/prof_match.py --rows 4000000 --dim 128 --metric l2 --batch 64 --iters 6 >> gpurun_out/r2s_l2.txt 2>&1
  timeout 300 python tools/prof_match.py --rows 4000000 --dim 128 --metric l2 --batch 1024 --iters 6 >> gpurun_out/r2s_l2.txt 2>&1
done
grep -v "^$" gpurun_out/r2s_l2.txt
: > gpurun_out/r2s_c1.jsonl
for c in 0 1; do
  FRP_SMALL_CLUSTER=$c timeout 200 python tools/prof_c1.py >> gpurun_out/r2s_c1.jsonl 2>gpurun_out/r2s_c1_$c.err || echo "c1 $c failed"
done
cat gpurun_out/r2s_c1.jsonl
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:match_small -c 30 --csv --log-file gpurun_out/r2s_small_sweep.csv python tools/prof_c1.py --sweep > gpurun_out/r2s_sweep.log 2>&1
grep -v "^==" gpurun_out/r2s_small_sweep.csv | awk -F'","' 'NR>1{print $8, $NF}' | tr -d '"' | paste - - - - 
timeout 900 python -m pytest tests/test_match_gpu.py tests/test_round2_gpu.py tests/test_shard_gpu.py -m gpu -x -q -k "not two_devices" 2>&1 | tail -3
