#!/bin/bash
# final single-GPU record after the issuer / small-gallery changes: smoke, GPU tests, default bench,
# then the ncu captures of the kernels whose source changed
mkdir -p gpurun_out
bash tools/gpu_run_final1.sh
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:match_small -c 30 --csv --log-file gpurun_out/r2_launches_small_sweep.csv python tools/prof_c1.py --sweep > gpurun_out/r2_small_sweep.log 2>&1
grep -v "^==" gpurun_out/r2_launches_small_sweep.csv | awk -F'","' 'NR>1{print $8, $NF}' | tr -d '"' | paste - - - -
NCU="ncu --set full --clock-control none --import-source on -f"
PM="python tools/prof_match.py"
run() { # name, kernel regex, skip, args...
  name=$1; re=$2; skip=$3; shift 3
  timeout 300 $PM "$@" --iters 1 > gpurun_out/ncu_plain_$name.log 2>&1 || { echo "$name: plain run failed"; tail -3 gpurun_out/ncu_plain_$name.log; return; }
  timeout 900 $NCU -k regex:$re -s $skip -c 1 -o gpurun_out/r2_$name $PM "$@" --iters 1 > gpurun_out/ncu_$name.log 2>&1
  echo "$name rc=$? $(tail -1 gpurun_out/ncu_plain_$name.log)"
}
run gemm_pair_b256_10m match_gemm_pair 0 --rows 10000000 --dim 512 --metric cos --batch 256
run gemm_l2_128_b256_4m match_gemm 0 --rows 4000000 --dim 128 --metric l2 --batch 256
run gemm_b1_10m match_gemm 0 --rows 10000000 --dim 512 --metric cos --batch 1
