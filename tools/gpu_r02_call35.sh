#!/bin/bash
# Round 2, GPU call 35 (1 GPU): isolated kernel durations of small G1 MSMs (2^14, 2^16): where do 7.9 / 3.35 ms go?
set -u
mkdir -p gpurun_out
for l in 14 16; do
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r02_c35_launches_g1_n$l.csv python bench.py --workload g1_msm --log2n $l --steps 1 --warmup 1 --no-cpu-baseline --no-e2e > gpurun_out/r02_c35_$l.log 2>&1
done
python - <<'PY'
import csv
for l in (14,16):
    rows=[r for r in csv.reader(open('gpurun_out/r02_c35_launches_g1_n%d.csv'%l)) if len(r)>10 and r[0].isdigit()]
    seq=[]
    for r in rows:
        name=r[4].split('(')[0].split('::')[-1][:30]
        val=float(r[-1].replace(',','')); unit=r[-2]
        if unit=='us': val/=1e3
        elif unit=='ns': val/=1e6
        seq.append((name,val,r[6] if len(r)>6 else ''))
    idx=[i for i,(n,v,g) in enumerate(seq) if n.startswith('k_msm_count')]
    # last MSM = last 4 counts
    start=idx[-4] if len(idx)>=4 else idx[0]
    print('n = 2^%d'%l)
    for n,v,g in seq[start:]: print('   %-32s %.3f'%(n,v))
PY
