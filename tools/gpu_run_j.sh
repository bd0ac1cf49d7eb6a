#!/bin/bash
mkdir -p gpurun_out
N=${1:-8}
i=0
for cfg in "0 " "2 " "0 --no-pipeline" "2 " "0 "; do
  i=$((i+1)); set -- $cfg; R=$1; shift
  FRP_RESERVED_SMS=$R timeout 240 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 2958$i bench.py --gpus $N --steps 60 --warmup 5 --no-extras --no-cpu "$@" > gpurun_out/r2j_n${N}_c$i.json 2> gpurun_out/r2j_n${N}_c$i.err
  python - <<PY
import json
try:
    d=json.loads([l for l in open('gpurun_out/r2j_n${N}_c$i.json') if l.startswith('{')][-1])
    print('reserve $R $@', d.get('stepping'),'value',round(d['value']),'ms',round(d['ms_per_step'],4),'ksum',round(sum(d['roofline']['kernel_ms'].values()),4),'e2e',round(d['e2e']['value']))
except Exception as e: print('cfg $i ERR',e)
PY
done
