#!/bin/bash
# Round 2, GPU call 7 (1 GPU): full GPU suite after the dispatch fix, default bench line, h2c bench, ncu of k_imad_peak
set -u
mkdir -p gpurun_out
python -m pytest tests -m gpu -q -p no:cacheprovider 2>&1 | tail -8 > gpurun_out/r02_c7_pytest.txt
python bench.py > gpurun_out/r02_c7_bench_all.json 2> gpurun_out/r02_c7_bench_all.err; echo "bench rc=$?" >> gpurun_out/r02_c7_bench_all.err
python tools/bench_h2c.py --log-n 16 > gpurun_out/r02_c7_h2c.json 2> gpurun_out/r02_c7_h2c.err
timeout 600 ncu --set full --clock-control none -k regex:k_imad_peak -c 4 -o gpurun_out/r02_ncu_imad_peak python -c "
import bls12_381_b200 as b
e=b.Engine(); print(e.imad_peak(2000), e.imad_peak(2000, mode=1)); e.close()" > gpurun_out/r02_c7_imad.txt 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_msm_accumulate -s 3 -c 1 -o gpurun_out/r02_ncu_g1acc \
    python bench.py --workload g1_msm --steps 1 --warmup 3 --no-cpu-baseline --no-e2e > /dev/null 2>&1
python __graft_entry__.py --smoke > gpurun_out/r02_c7_smoke.txt 2>&1
cat gpurun_out/r02_c7_pytest.txt; tail -3 gpurun_out/r02_c7_bench_all.err; cat gpurun_out/r02_c7_h2c.json | cut -c1-700; tail -2 gpurun_out/r02_c7_imad.txt gpurun_out/r02_c7_smoke.txt
python - <<'PY'
import json
d=json.loads(open('gpurun_out/r02_c7_bench_all.json').read().strip().splitlines()[-1])
print('headline', d['ms_per_step'], d['value'], d['e2e']['ms_per_step'], d['roofline']['frac'], d['roofline'].get('peak_detail'))
print(d['cpu_baseline'])
for k,c in d.get('configs',{}).items():
    print(k, c['ms_per_step'], c['value'], (c['e2e'] or {}).get('ms_per_step'), (c['roofline'] or {}).get('frac'), (c['roofline'] or {}).get('model_frac_whole_step'), (c['cpu_baseline'] or {}).get('value'))
PY
