#!/bin/bash
# MMA issuer: divergent `if (lane == 0)` vs warp-convergent elect.sync issue (A/B), then the tests
mkdir -p gpurun_out
LIB=face-recognition-platform_b200/frp_b200/_lib/libfrp_b200.so
cp $LIB /tmp/default.so
: > gpurun_out/r2r_elect.txt
for v in elect0 elect1 elect0 elect1; do
  cp tools/variants/$v.so $LIB
  echo "== $v" >> gpurun_out/r2r_elect.txt
  timeout 300 python tools/prof_match.py --rows 4000000 --dim 128 --metric l2 --batch 256 --iters 6 >> gpurun_out/r2r_elect.txt 2>&1
  timeout 300 python tools/prof_match.py --rows 4000000 --dim 128 --metric l2 --batch 64 --iters 6 >> gpurun_out/r2r_elect.txt 2>&1
  timeout 300 python tools/prof_match.py --rows 4000000 --dim 128 --metric cos --batch 1024 --iters 6 >> gpurun_out/r2r_elect.txt 2>&1
  timeout 300 python tools/prof_match.py --rows 10000000 --dim 512 --metric cos --batch 256 --iters 6 >> gpurun_out/r2r_elect.txt 2>&1
  timeout 300 python tools/prof_match.py --rows 10000000 --dim 512 --metric cos --batch 4 --iters 6 >> gpurun_out/r2r_elect.txt 2>&1
done
cp /tmp/default.so $LIB
grep -v "^$" gpurun_out/r2r_elect.txt
timeout 900 python -m pytest tests/test_match_gpu.py tests/test_round2_gpu.py -m gpu -x -q -k "not two_devices" 2>&1 | tail -3
