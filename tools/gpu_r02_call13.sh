#!/bin/bash
# Round 2, GPU call 13 (1 GPU): isolated kernel durations (ncu serialises launches) of the G1 MSM with both counting sorts
set -u
mkdir -p gpurun_out
for srt in 1 0; do
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r02_c13_launches_sort$srt.csv python bench.py --workload g1_msm --steps 1 --warmup 1 --no-cpu-baseline --no-e2e --tune msm_sort=$srt > gpurun_out/r02_c13_b$srt.log 2>&1
done
python - <<'PY'
import csv,collections
for srt in (1,0):
    rows=[r for r in csv.reader(open('gpurun_out/r02_c13_launches_sort%d.csv'%srt)) if len(r)>10 and r[0].isdigit()]
    acc=collections.OrderedDict()
    for r in rows:
        name=r[4].split('(')[0][-40:]
        val=float(r[-1].replace(',',''))
        unit=r[-2]
        if unit=='us': val/=1e3
        elif unit=='ns': val/=1e6
        elif unit=='s' or unit=='second': val*=1e3
        acc.setdefault(name,[]).append(val)
    print('sort',srt)
    for k,v in acc.items(): print('   %-42s n=%3d sum=%.3f ms  each=%s'%(k,len(v),sum(v),[round(x,3) for x in v[-8:]]))
PY
