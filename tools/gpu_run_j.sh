#!/bin/bash
mkdir -p gpurun_out
N=${1:-2}
i=0
for mode in "" "--no-graph" "--pipeline"; do
  i=$((i+1))
  timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 2954$i bench.py --gpus $N --steps 40 --warmup 5 --no-extras --no-cpu $mode > gpurun_out/r2j_n${N}_m$i.json 2> gpurun_out/r2j_n${N}_m$i.err
  echo "mode '$mode' rc=$?"
  grep -i "capture failed\|Traceback\|Error" gpurun_out/r2j_n${N}_m$i.err | head -3
done
python - <<'PY'
import json,glob
for f in sorted(glob.glob('gpurun_out/r2j_n*_m*.json')):
    try:
        d=json.loads([l for l in open(f) if l.startswith('{')][-1])
        print(f, d.get('stepping'),'value',round(d['value']),'ms',round(d['ms_per_step'],4),'ksum',round(sum(d['roofline']['kernel_ms'].values()),4),'e2e',round(d['e2e']['value']),'launches',d['gpu_launches'])
    except Exception as e: print(f,'ERR',e)
PY
