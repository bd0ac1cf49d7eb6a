#!/bin/bash
# final single-GPU record: smoke, GPU tests, default bench line (reference arm: see r2_bench_reference_n1.json)
mkdir -p gpurun_out
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/r2_pytest_1gpu.log 2>&1
echo "pytest rc=$?"; tail -3 gpurun_out/r2_pytest_1gpu.log
timeout 900 python bench.py > gpurun_out/r2_bench_n1.json 2> gpurun_out/r2_bench_n1.err
echo "bench rc=$?"; tail -2 gpurun_out/r2_bench_n1.err
python - <<'PY'
import json
d=json.load(open('gpurun_out/r2_bench_n1.json'))
print('value',d['value'],'e2e',d['e2e']['value'],'roof',d['roofline']['frac'], d['roofline']['kernel_ms'], d.get('stepping'))
x=d['extra']
print('c1',{k:v for k,v in x['c1_compare_faces'].items() if k.endswith('_us') or k.endswith('_ms')})
print('align',x['align']['kernel_ms'],x['align']['call_ms'],x['align']['roofline']['frac'])
for k in ('match_batch1','match_topk_nothreshold','c2_1m','cluster_faces','pipeline_c5'): print(k, json.dumps(x[k])[:400])
print('l2', json.dumps(x['match_l2_128d']['batches'])[:600])
PY
