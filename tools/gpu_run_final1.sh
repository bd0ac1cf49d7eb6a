#!/bin/bash
# final single-GPU record: smoke, GPU tests, default bench line, reference arm
mkdir -p gpurun_out
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -3
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/r2_pytest_1gpu.log 2>&1
echo "pytest rc=$?"; tail -3 gpurun_out/r2_pytest_1gpu.log
timeout 900 python bench.py > gpurun_out/r2_bench_n1.json 2> gpurun_out/r2_bench_n1.err
echo "bench rc=$?"; tail -2 gpurun_out/r2_bench_n1.err
( time timeout 600 python bench.py --impl reference > gpurun_out/r2_bench_reference.json 2> gpurun_out/r2_bench_reference.err ) 2>&1 | grep real
python - <<'PY'
import json
d=json.load(open('gpurun_out/r2_bench_n1.json'))
print('value',d['value'],'e2e',d['e2e']['value'],'roof',d['roofline']['frac'], d['roofline']['kernel_ms'], d.get('stepping'), d['clocks'])
x=d['extra']
print('c1',{k:v for k,v in x['c1_compare_faces'].items() if k.endswith('_us') or k.endswith('_ms')})
print('align',x['align']['kernel_ms'],x['align']['call_ms'],x['align']['roofline']['frac'])
r=json.load(open('gpurun_out/r2_bench_reference.json'))
print('reference',r['value'],r['steps'],r['ms_per_step'],r['cpu_baseline']['cores'])
PY
