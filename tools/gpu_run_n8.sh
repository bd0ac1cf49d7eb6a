#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29521 bench.py --gpus 8 --steps 30 --warmup 3 > gpurun_out/r2_bench_n8.json 2> gpurun_out/r2_bench_n8.err
echo "bench n8 rc=$?"
grep -v "^\[W\|NCCL\|^W1\|^\*\*\*\|OMP_NUM" gpurun_out/r2_bench_n8.err | tail -8
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29522 bench.py --gpus 8 --steps 30 --warmup 3 --no-pipeline --no-extras > gpurun_out/r2_bench_n8_nopipe.json 2> gpurun_out/r2_bench_n8_nopipe.err
echo "bench n8 nopipe rc=$?"
python - <<'PY'
import json
for f in ('gpurun_out/r2_bench_n8.json','gpurun_out/r2_bench_n8_nopipe.json'):
    try:
        d=json.loads([l for l in open(f) if l.startswith('{')][-1])
        print(f, 'value',round(d['value']),'ms',d['ms_per_step'],'e2e',round(d['e2e']['value']),'share',d['roofline']['kernel_share_of_step'], d['roofline']['kernel_ms'])
        if 'extra' in d:
            for k,v in d['extra'].items(): print('  ',k, json.dumps(v)[:1200])
    except Exception as e: print(f,'ERR',e)
PY
