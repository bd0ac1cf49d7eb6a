#!/bin/bash
# 2-GPU validation: pipelined bench at N=2 (+ unpipelined for comparison) and the sampled-cut test
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_round2_gpu.py -m gpu -x -q -k "pure_topk" 2>&1 | tail -2
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 20 --warmup 3 > gpurun_out/r2b_bench_n2.json 2> gpurun_out/r2b_bench_n2.err
echo "bench n2 rc=$?"
grep -v "^\[W\|NCCL\|^W1\|^\*\*\*\|OMP_NUM" gpurun_out/r2b_bench_n2.err | tail -8
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 2 --steps 20 --warmup 3 --no-pipeline --no-extras > gpurun_out/r2b_bench_n2_nopipe.json 2> gpurun_out/r2b_bench_n2_nopipe.err
echo "bench n2 nopipe rc=$?"
python - <<'PY'
import json
for f in ('gpurun_out/r2b_bench_n2.json','gpurun_out/r2b_bench_n2_nopipe.json'):
    try:
        d=json.loads([l for l in open(f) if l.startswith('{')][-1])
        print(f, 'value',round(d['value']),'ms',d['ms_per_step'],'e2e',round(d['e2e']['value']),'share',d['roofline']['kernel_share_of_step'], d['roofline']['kernel_ms'])
        if 'extra' in d:
            for k,v in d['extra'].items(): print('  ',k, json.dumps(v)[:300])
    except Exception as e: print(f,'ERR',e)
PY
