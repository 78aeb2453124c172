#!/bin/bash
# Round 2, GPU call 3: six-lane pairing kernels with called (non-inlined) group operations + padded board; new bench.py
set -u
mkdir -p gpurun_out
python -m pytest tests -m gpu -q -x -p no:cacheprovider 2>&1 | tail -5 > gpurun_out/r02_c3_pytest.txt
for w in 8 12 16; do
python bench.py --workload pairing --steps 5 --warmup 3 --no-cpu-baseline --no-e2e --tune coop_warps=$w > gpurun_out/r02_c3_pairing_w$w.json 2>> gpurun_out/r02_c3_pairing.err
done
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_coop_pairing -s 3 -c 1 -o gpurun_out/r02_ncu_coop2_w12 \
    python bench.py --workload pairing --steps 1 --warmup 3 --no-cpu-baseline --no-e2e --tune coop_warps=12 > /dev/null 2>&1
python bench.py > gpurun_out/r02_c3_bench_all.json 2> gpurun_out/r02_c3_bench_all.err; echo "bench rc=$?" >> gpurun_out/r02_c3_bench_all.err
python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/r02_c3_bench_ref.json 2> gpurun_out/r02_c3_bench_ref.err
cat gpurun_out/r02_c3_pytest.txt
for f in gpurun_out/r02_c3_pairing_w*.json; do echo $f; python -c "
import json,sys
d=json.load(open('$f')); r=d.get('roofline') or {}
print(d['ms_per_step'], d['value'], r.get('frac'), r.get('kernel_ms'))"; done
tail -5 gpurun_out/r02_c3_pairing.err gpurun_out/r02_c3_bench_all.err
python - <<'PY'
import json
d=json.load(open('gpurun_out/r02_c3_bench_all.json'))
print('headline', d['ms_per_step'], d['value'], d['e2e']['ms_per_step'], d['roofline']['frac'], d['cpu_baseline']['value'], d['cpu_baseline']['single_thread'], d['cpu_baseline']['effective_cores'], d['cpu_baseline']['host'])
for k,c in d.get('configs',{}).items():
    print(k, c['ms_per_step'], c['value'], (c['e2e'] or {}).get('ms_per_step'), (c['roofline'] or {}).get('frac'), (c['roofline'] or {}).get('model_frac_whole_step'), (c['cpu_baseline'] or {}).get('value'))
PY
