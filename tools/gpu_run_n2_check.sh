#!/bin/bash
# 2-GPU sanity of the final build: the two-rank NCCL / p2p test, the two-device test, a short sharded bench
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_round2_gpu.py tests/test_shard_gpu.py -m gpu -q -k "two_rank or two_devices or nccl or sharded" > gpurun_out/r2_pytest_2gpu.log 2>&1
echo "pytest rc=$?"; tail -3 gpurun_out/r2_pytest_2gpu.log
timeout 420 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29562 bench.py --gpus 2 --steps 30 --warmup 3 > gpurun_out/r2_bench_n2.json 2> gpurun_out/r2_bench_n2.err
echo "bench n2 rc=$?"
grep -i "Traceback\|Error" gpurun_out/r2_bench_n2.err | head -3
python - <<PY
import json
d=json.loads([l for l in open('gpurun_out/r2_bench_n2.json') if l.startswith('{')][-1])
print('N=2', d.get('stepping'),'value',round(d['value']),'ms',round(d['ms_per_step'],4),'e2e',round(d['e2e']['value']),'kernels',d['roofline']['kernel_ms'])
for k,v in d.get('extra',{}).items(): print('  ',k, json.dumps(v)[:400])
PY
