#!/bin/bash
mkdir -p gpurun_out
timeout 300 python tools/prof_step.py > gpurun_out/plain_step.log 2>&1 && \
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"matvec_kernel|add_rmsnorm|attn_decode|swiglu|argmax|embed" -s 260 -c 520 --csv --log-file gpurun_out/launches_step_r1.csv python tools/prof_step.py > gpurun_out/ncu_step.log 2>&1
tail -2 gpurun_out/ncu_step.log
python - <<'PY'
import csv, collections
rows=[r for r in csv.reader(open('gpurun_out/launches_step_r1.csv')) if len(r)>5]
hdr=rows[0]; ik=hdr.index('Kernel Name'); iv=hdr.index('Metric Value')
agg=collections.defaultdict(lambda:[0,0.0])
for r in rows[1:]:
    k=r[ik][:70]; agg[k][0]+=1; agg[k][1]+=float(r[iv].replace(',',''))
tot=sum(v[1] for v in agg.values())
for k,v in sorted(agg.items(), key=lambda kv:-kv[1][1]):
    print(f"{k:72s} n={v[0]:4d} total={v[1]/1e3:9.1f} us  avg={v[1]/v[0]/1e3:7.2f} us  share={100*v[1]/tot:5.1f}%")
print("total us (2 steps)", tot/1e3)
PY
