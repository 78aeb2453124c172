#!/bin/bash
# Round 2, GPU call 34 (1 GPU): final state — whole -m gpu suite, smoke, the default bench line, launch list of the same command
set -u
mkdir -p gpurun_out
python -m pytest tests -m gpu -q -x -p no:cacheprovider 2>&1 | tail -4 > gpurun_out/r02_c34_pytest.txt
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r02_c34_smoke.txt 2>&1
python bench.py > gpurun_out/r02_c34_bench_all.json 2> gpurun_out/r02_c34_bench_all.err; echo "rc=$?" >> gpurun_out/r02_c34_bench_all.err
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r02_c34_launches_default_bench.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-e2e > gpurun_out/r02_c34_ncu.log 2>&1
cat gpurun_out/r02_c34_pytest.txt; tail -1 gpurun_out/r02_c34_smoke.txt
python - <<'PY'
import json
d=json.loads(open('gpurun_out/r02_c34_bench_all.json').read().strip().splitlines()[-1])
def show(name,c):
    r=c.get('roofline') or {}
    print(name,'ms',round(c['ms_per_step'],3),'value','%.4g'%c['value'],'e2e',round(c['e2e']['ms_per_step'],3) if c.get('e2e') else None,'frac',round(r.get('frac',0),3),'exec',round(r.get('executed_frac',0),3),'whole',round(r.get('model_frac_whole_step',0),3), 'traffic', r.get('traffic'))
    cb=c.get('cpu_baseline')
    if cb: print('    cpu','%.4g'%cb['value'],cb['cores'],'single','%.4g'%cb['single_thread']['value'])
show('g1_msm',d)
for k,c in d['configs'].items(): show(k,c)
print(d['clocks'], d['gpu_launches'])
PY
tail -n 2 gpurun_out/r02_c34_bench_all.err
