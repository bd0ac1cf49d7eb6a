#!/bin/bash
# alignment fix check + epilogue A/B (parked candidates vs insert-on-the-spot) + full GPU tests
mkdir -p gpurun_out
LIB=face-recognition-platform_b200/frp_b200/_lib/libfrp_b200.so
for st in "2.8 0.25" "1.9 0.25"; do
  timeout 120 python tools/align_bisect.py one $st 2>&1 | grep -v "^$" | tail -2
done
timeout 600 python -m pytest tests/test_align_gpu.py -m gpu -x -q 2>&1 | tail -3
timeout 300 python tools/prof_align.py v6fix 2>&1 | tail -1
timeout 300 python tools/prof_align.py v6fix 2>&1 | tail -1
echo "== epilogue: new (parked candidates)"
timeout 300 python tools/prof_match.py --rows 10000000 --dim 512 --metric cos --batch 256 --flags 0x1 --iters 4 2>&1 | tail -1
timeout 300 python tools/prof_match.py --rows 10000000 --dim 512 --metric cos --batch 256 --iters 4 2>&1 | tail -1
timeout 300 python tools/prof_match.py --rows 4000000 --dim 128 --metric l2 --batch 256 --iters 4 2>&1 | tail -1
timeout 300 python tools/prof_match.py --rows 4000000 --dim 128 --metric l2 --batch 256 --flags 0x1 --iters 4 2>&1 | tail -1
cp $LIB /tmp/default.so
cp tools/variants/v6.so $LIB
echo "== epilogue: old (insert on the spot)"
timeout 300 python tools/prof_match.py --rows 10000000 --dim 512 --metric cos --batch 256 --flags 0x1 --iters 4 2>&1 | tail -1
timeout 300 python tools/prof_match.py --rows 10000000 --dim 512 --metric cos --batch 256 --iters 4 2>&1 | tail -1
timeout 300 python tools/prof_match.py --rows 4000000 --dim 128 --metric l2 --batch 256 --iters 4 2>&1 | tail -1
timeout 300 python tools/prof_match.py --rows 4000000 --dim 128 --metric l2 --batch 256 --flags 0x1 --iters 4 2>&1 | tail -1
cp /tmp/default.so $LIB
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/r2f_pytest.log 2>&1
echo "pytest rc=$?"; tail -4 gpurun_out/r2f_pytest.log
