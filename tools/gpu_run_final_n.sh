#!/bin/bash
# final multi-GPU records: bash tools/gpu_run_final_n.sh N [tests]
N=$1
mkdir -p gpurun_out
if [ "$2" = "tests" ]; then
  timeout 900 python -m pytest tests/test_round2_gpu.py tests/test_shard_gpu.py -m gpu -q > gpurun_out/r2_pytest_${N}gpu.log 2>&1
  echo "pytest rc=$?"; tail -3 gpurun_out/r2_pytest_${N}gpu.log
fi
timeout 420 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 2956$N bench.py --gpus $N --steps 30 --warmup 3 > gpurun_out/r2_bench_n$N.json 2> gpurun_out/r2_bench_n$N.err
echo "bench n$N rc=$?"
grep -i "Traceback\|Error" gpurun_out/r2_bench_n$N.err | head -3
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 2957$N bench.py --impl reference --gpus $N --steps 3 --warmup 1 > gpurun_out/r2_bench_reference_n$N.json 2> gpurun_out/r2_bench_reference_n$N.err
python - <<PY
import json
d=json.loads([l for l in open('gpurun_out/r2_bench_n$N.json') if l.startswith('{')][-1])
print('N=$N', d.get('stepping'),'value',round(d['value']),'ms',round(d['ms_per_step'],4),'e2e',round(d['e2e']['value']),'kernels',d['roofline']['kernel_ms'])
for k,v in d.get('extra',{}).items(): print('  ',k, json.dumps(v)[:700])
r=json.loads([l for l in open('gpurun_out/r2_bench_reference_n$N.json') if l.startswith('{')][-1])
print('reference at N=$N:', r['value'], 'cores', r['cpu_baseline']['cores'], 'steps', r['steps'])
PY
