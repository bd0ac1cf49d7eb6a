#!/bin/bash
mkdir -p gpurun_out
for tag in a b; do
timeout 600 python bench.py --steps 40 --warmup 5 --no-extras --no-cpu > gpurun_out/r2k_$tag.json 2> gpurun_out/r2k_$tag.err
done
timeout 600 python bench.py --steps 40 --warmup 5 --no-extras --no-cpu --no-pipeline > gpurun_out/r2k_nopipe.json 2> gpurun_out/r2k_nopipe.err
python - <<'PY'
import json,glob
for f in sorted(glob.glob('gpurun_out/r2k_*.json')):
    try:
        d=json.loads([l for l in open(f) if l.startswith('{')][-1])
        km=d['roofline']['kernel_ms']
        print(f, 'value',round(d['value']),'ms',round(d['ms_per_step'],4),'kernel sum',round(sum(km.values()),4),'e2e',round(d['e2e']['value']), d['e2e'].get('ms_per_step'))
    except Exception as e: print(f,'ERR',e)
PY
