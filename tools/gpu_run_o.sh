#!/bin/bash
for lpc in 2 4 2 4; do
  echo "LPC=$lpc"
  FRP_GEMM_LPC=$lpc timeout 200 python tools/prof_match.py --rows 4000000 --dim 128 --metric l2 --batch 256 --iters 5 2>&1 | tail -1
  FRP_GEMM_LPC=$lpc timeout 200 python tools/prof_match.py --rows 4000000 --dim 128 --metric cos --batch 256 --iters 5 2>&1 | tail -1
done
FRP_GEMM_LPC=4 timeout 600 python -m pytest tests/test_match_gpu.py -m gpu -x -q -k "tcgen05_path_other_dims or dims_and_metrics or exact_ties" 2>&1 | tail -2
