#!/bin/bash
# 128-d rows: one-CTA kernel with 2-CTA multicast (default) vs the cta_group::2 pair kernel, after the issuer change
mkdir -p gpurun_out
: > gpurun_out/r2v_pair128.txt
for rep in 1 2; do for pair in 0 1; do
  echo "== FRP_GEMM_PAIR=$pair" >> gpurun_out/r2v_pair128.txt
  FRP_GEMM_PAIR=$pair timeout 300 python tools/prof_match.py --rows 4000000 --dim 128 --metric l2 --batch 256 --iters 6 >> gpurun_out/r2v_pair128.txt 2>&1
  FRP_GEMM_PAIR=$pair timeout 300 python tools/prof_match.py --rows 4000000 --dim 128 --metric cos --batch 256 --iters 6 >> gpurun_out/r2v_pair128.txt 2>&1
  FRP_GEMM_PAIR=$pair timeout 300 python tools/prof_match.py --rows 4000000 --dim 128 --metric l2 --batch 1024 --iters 6 >> gpurun_out/r2v_pair128.txt 2>&1
done; done
grep -v "^$" gpurun_out/r2v_pair128.txt
