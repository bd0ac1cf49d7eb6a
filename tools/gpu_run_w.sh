#!/bin/bash
# programmatic dependent launches prep -> prefilter -> re-rank: A/B on the default bench, then the tests
mkdir -p gpurun_out
for p in 0 1 0 1; do
  FRP_PDL=$p timeout 600 python bench.py --no-cpu > gpurun_out/r2w_bench_pdl$p.json 2> gpurun_out/r2w_bench_pdl$p.err || echo "bench pdl=$p failed"
  python - <<PY
import json
d=json.loads([l for l in open('gpurun_out/r2w_bench_pdl$p.json') if l.startswith('{')][-1])
x=d['extra']
print('PDL=$p value',round(d['value']),'ms',round(d['ms_per_step'],4),'e2e',round(d['e2e']['value']),'c2',round(x['c2_1m']['queries_per_s']),round(x['c2_1m']['ms_per_step'],4),'c2 e2e',round(x['c2_1m']['e2e_queries_per_s']),'nothr',round(x['match_topk_nothreshold']['kernel_ms'],3),'c5',round(x['pipeline_c5']['ms_per_tick'],3))
PY
done
timeout 1200 python -m pytest tests -m gpu -x -q 2>&1 | tail -3
