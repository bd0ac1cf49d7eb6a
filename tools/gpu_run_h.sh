#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_round2_gpu.py tests/test_align_gpu.py -m gpu -x -q -k "pure_topk or align or large_topk" 2>&1 | tail -4
timeout 300 python tools/prof_align.py v6d 2>&1 | tail -1
timeout 900 python bench.py --steps 20 --warmup 3 > gpurun_out/r2h_bench.json 2> gpurun_out/r2h_bench.err
echo "bench rc=$?"; tail -3 gpurun_out/r2h_bench.err
python - <<'PY'
import json
d=json.load(open('gpurun_out/r2h_bench.json'))
print('value',d['value'],'e2e',d['e2e']['value'],'roof',d['roofline']['frac'], d['roofline']['kernel_ms'])
x=d['extra']
print('nothr',x['match_topk_nothreshold'])
print('align',x['align']['kernel_ms'],x['align']['roofline']['frac'])
print('c1',x['c1_compare_faces']['frp_match_call_us'],x['c1_compare_faces']['store_face_us'],x['c1_compare_faces']['match_topk_ms'])
print('l2',x['match_l2_128d']['batches']['256'])
PY
