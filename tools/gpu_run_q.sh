#!/bin/bash
# round-2 late experiments: TMEM read probe, small-gallery latency switches, epilogue prefetch A/B
mkdir -p gpurun_out
LIB=face-recognition-platform_b200/frp_b200/_lib/libfrp_b200.so
cp $LIB /tmp/default.so
timeout 120 tools/variants/tmem_ld_bench > gpurun_out/r2_tmem_ld_probe.jsonl 2>&1; echo "probe rc=$?"
cat gpurun_out/r2_tmem_ld_probe.jsonl
: > gpurun_out/r2q_c1.jsonl
for c in 0 1; do for p in 0 1; do
  FRP_SMALL_CLUSTER=$c FRP_SMALL_POLL=$p timeout 200 python tools/prof_c1.py >> gpurun_out/r2q_c1.jsonl 2>gpurun_out/r2q_c1_$c$p.err || echo "c1 $c$p failed"
done; done
cat gpurun_out/r2q_c1.jsonl
for c in 0 1; do
  FRP_SMALL_CLUSTER=$c timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:match_small -s 40 -c 8 --csv --log-file gpurun_out/r2q_small_$c.csv python tools/prof_c1.py > /dev/null 2>&1
  grep -v "^==" gpurun_out/r2q_small_$c.csv | tail -4 | awk -F'","' '{print $5, $8, $NF}'
done
: > gpurun_out/r2q_prefetch.txt
for v in nopref pref nopref pref; do
  cp tools/variants/$v.so $LIB
  echo "== $v" >> gpurun_out/r2q_prefetch.txt
  timeout 300 python tools/prof_match.py --rows 4000000 --dim 128 --metric l2 --batch 256 --iters 6 >> gpurun_out/r2q_prefetch.txt 2>&1
  timeout 300 python tools/prof_match.py --rows 4000000 --dim 128 --metric l2 --batch 64 --iters 6 >> gpurun_out/r2q_prefetch.txt 2>&1
  timeout 300 python tools/prof_match.py --rows 10000000 --dim 512 --metric cos --batch 256 --iters 6 >> gpurun_out/r2q_prefetch.txt 2>&1
done
cp /tmp/default.so $LIB
grep -v "^$" gpurun_out/r2q_prefetch.txt
timeout 900 python -m pytest tests/test_match_gpu.py tests/test_round2_gpu.py -m gpu -x -q -k "not two_devices" 2>&1 | tail -3
