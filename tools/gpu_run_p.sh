#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_round2_gpu.py -m gpu -x -q -k "batching or small_galleries or cooldown or cluster" 2>&1 | tail -2
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > /dev/null 2>&1 && \
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 300 --csv --log-file gpurun_out/r2_launches_smoke.csv python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r2p_ncu.log 2>&1
echo "ncu rc=$?"
grep -v "^==" gpurun_out/r2_launches_smoke.csv | python -c "
import csv,sys,collections
rows=list(csv.reader(sys.stdin)); h=rows[0]; ik=h.index('Kernel Name'); iv=h.index('Metric Value')
t=collections.OrderedDict()
for r in rows[1:]:
    if len(r)>iv:
        k=r[ik].split('(')[0][-44:]; t.setdefault(k,[]).append(float(r[iv].replace(',',''))/1000)
for k,v in t.items(): print('%-46s n=%3d  mean %9.1f us'%(k,len(v),sum(v)/len(v)))
"
