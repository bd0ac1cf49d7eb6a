#!/bin/bash
# alignment kernel variants: correctness (bit-exact tests) + C4 timing, then one ncu capture of the default
mkdir -p gpurun_out
LIB=face-recognition-platform_b200/frp_b200/_lib/libfrp_b200.so
cp $LIB /tmp/default.so
: > gpurun_out/r2c_align_variants.jsonl
for v in v5 v6 v6nopred v6w21 v6w16; do
  cp tools/variants/$v.so $LIB
  if [ "$v" != "v5" ]; then
    timeout 600 python -m pytest tests/test_align_gpu.py -m gpu -x -q > gpurun_out/r2c_pytest_$v.log 2>&1
    echo "$v pytest rc=$?"; tail -3 gpurun_out/r2c_pytest_$v.log
  fi
  timeout 300 python tools/prof_align.py $v >> gpurun_out/r2c_align_variants.jsonl 2>gpurun_out/r2c_prof_$v.err
  timeout 300 python tools/prof_align.py $v >> gpurun_out/r2c_align_variants.jsonl 2>>gpurun_out/r2c_prof_$v.err
done
cat gpurun_out/r2c_align_variants.jsonl
cp /tmp/default.so $LIB
timeout 600 ncu --set full --clock-control none --import-source on -k regex:align_kernel -c 1 -f -o gpurun_out/r2_align_v6 python tools/prof_align.py --once > gpurun_out/r2c_ncu.log 2>&1
echo "ncu rc=$?"; tail -3 gpurun_out/r2c_ncu.log
