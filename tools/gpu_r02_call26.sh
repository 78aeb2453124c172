#!/bin/bash
# Round 2, GPU call 26 (1 GPU): isolated kernel durations of ONE window shard (shard 7 of 8) of the 2^20 G1 MSM, one group / one window per group
set -u
mkdir -p gpurun_out
for tg in 0 1; do
ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file gpurun_out/r02_c26_shard7_tg$tg.csv python tools/shard_once.py 7 8 msm_tail_groups=$tg msm_coop_nloc=3 > gpurun_out/r02_c26_$tg.log 2>&1
done
python - <<'PY'
import csv
for tg in (0,1):
    rows=[r for r in csv.reader(open('gpurun_out/r02_c26_shard7_tg%d.csv'%tg)) if len(r)>10 and r[0].isdigit()]
    print('tail_groups',tg)
    seq=[]
    for r in rows:
        name=r[4].split('(')[0].split('::')[-1][:28]
        val=float(r[-1].replace(',','')); unit=r[-2]
        if unit=='us': val/=1e3
        elif unit=='ns': val/=1e6
        seq.append((name,val))
    # last MSM only: find last k_msm_count index
    idx=[i for i,(n,v) in enumerate(seq) if n.startswith('k_msm_count')]
    ng = 2 if tg else 1
    start=idx[-ng]
    for n,v in seq[start:]: print('   %-30s %.3f'%(n,v))
PY
