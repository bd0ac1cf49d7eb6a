#!/bin/bash
mkdir -p gpurun_out
: > gpurun_out/r2x_pdl_steps.jsonl
for rep in 1 2 3; do for p in 0 1; do
  FRP_PDL=$p timeout 200 python tools/prof_step.py --rows 1000000 --dim 512 --batch 256 --steps 400 >> gpurun_out/r2x_pdl_steps.jsonl 2>gpurun_out/r2x.err
done; done
for p in 0 1; do
  FRP_PDL=$p timeout 200 python tools/prof_step.py --rows 1250000 --dim 512 --batch 256 --steps 400 >> gpurun_out/r2x_pdl_steps.jsonl 2>>gpurun_out/r2x.err
  FRP_PDL=$p timeout 200 python tools/prof_step.py --rows 500000 --dim 128 --batch 64 --steps 400 >> gpurun_out/r2x_pdl_steps.jsonl 2>>gpurun_out/r2x.err
done
cat gpurun_out/r2x_pdl_steps.jsonl
