#!/bin/bash
mkdir -p gpurun_out
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -8
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/r2l_pytest.log 2>&1
echo "pytest rc=$?"; tail -5 gpurun_out/r2l_pytest.log
timeout 900 python bench.py --steps 20 --warmup 3 > gpurun_out/r2l_bench.json 2> gpurun_out/r2l_bench.err
echo "bench rc=$?"; tail -3 gpurun_out/r2l_bench.err
python - <<'PY'
import json
d=json.load(open('gpurun_out/r2l_bench.json'))
print('value',d['value'],'e2e',d['e2e']['value'],'roof',d['roofline']['frac'], d['roofline']['kernel_ms'], d.get('stepping'))
x=d['extra']
print('c1',x['c1_compare_faces'])
print('align',x['align']['kernel_ms'],x['align']['call_ms'],x['align']['roofline']['frac'])
print('c5',x['pipeline_c5'])
PY
