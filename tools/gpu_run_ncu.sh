#!/bin/bash
# ncu evidence for the shipped kernels (each command has already exited 0 without ncu in this round)
mkdir -p gpurun_out
NCU="ncu --set full --clock-control none --import-source on -f"
PM="python tools/prof_match.py"
run() { # name, kernel regex, skip, args...
  name=$1; re=$2; skip=$3; shift 3
  timeout 300 $PM "$@" --iters 1 > gpurun_out/ncu_plain_$name.log 2>&1 || { echo "$name: plain run failed"; tail -3 gpurun_out/ncu_plain_$name.log; return; }
  timeout 900 $NCU -k regex:$re -s $skip -c 1 -o gpurun_out/r2_$name $PM "$@" --iters 1 > gpurun_out/ncu_$name.log 2>&1
  echo "$name rc=$? $(tail -1 gpurun_out/ncu_plain_$name.log)"
}
run gemm_pair_b256_10m match_gemm_pair 0 --rows 10000000 --dim 512 --metric cos --batch 256
run gemm_b1_10m match_gemm 0 --rows 10000000 --dim 512 --metric cos --batch 1
run gemm_nothr_b256_10m match_gemm_pair 1 --rows 10000000 --dim 512 --metric cos --batch 256 --flags 0x1
run gemm_l2_128_b256_4m match_gemm 0 --rows 4000000 --dim 128 --metric l2 --batch 256
run rerank_b256_10m rerank_kernel 0 --rows 10000000 --dim 512 --metric cos --batch 256
# launch lists of the default bench command (batch 256) and the batch-1 side leg
timeout 600 python bench.py --steps 2 --warmup 1 --no-extras --no-cpu > gpurun_out/ncu_plain_bench.log 2>&1 && \
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/r2_launches_batch256_bench.csv python bench.py --steps 2 --warmup 1 --no-extras --no-cpu > gpurun_out/ncu_bench.log 2>&1
echo "launch list rc=$?"
ls -la gpurun_out/*.ncu-rep | tail -8
