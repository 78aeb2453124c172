#!/bin/bash
# Round 2, GPU call 17 (1 GPU): isolated kernel durations (ncu serialises launches) of the G2 MSM and the pairing batch
set -u
mkdir -p gpurun_out
ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/r02_c17_launches_g2msm.csv python bench.py --workload g2_msm --steps 1 --warmup 1 --no-cpu-baseline --no-e2e > gpurun_out/r02_c17_g2.log 2>&1
ncu --metrics gpu__time_duration.sum --clock-control none -c 100 --csv --log-file gpurun_out/r02_c17_launches_pairing.csv python bench.py --workload pairing --steps 1 --warmup 1 --no-cpu-baseline --no-e2e > gpurun_out/r02_c17_pair.log 2>&1
python - <<'PY'
import csv,collections
for f in ('g2msm','pairing'):
    rows=[r for r in csv.reader(open('gpurun_out/r02_c17_launches_%s.csv'%f)) if len(r)>10 and r[0].isdigit()]
    acc=collections.OrderedDict()
    for r in rows:
        name=r[4].split('(')[0][-40:]
        val=float(r[-1].replace(',',''))
        unit=r[-2]
        if unit=='us': val/=1e3
        elif unit=='ns': val/=1e6
        elif unit in ('s','second'): val*=1e3
        acc.setdefault(name,[]).append(val)
    print(f)
    for k,v in acc.items(): print('   %-42s n=%3d sum=%.3f ms  each=%s'%(k,len(v),sum(v),[round(x,3) for x in v[-9:]]))
PY
