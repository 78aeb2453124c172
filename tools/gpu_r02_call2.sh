#!/bin/bash
# Round 2, GPU call 2: first hardware run of the six-lane pairing kernels (pairing_variant 7)
set -u
mkdir -p gpurun_out
python -m pytest tests -m gpu -q -x -p no:cacheprovider 2>&1 | tail -15 > gpurun_out/r02_c2_pytest.txt
for w in 8 12 16; do
python bench.py --workload pairing --steps 5 --warmup 3 --no-cpu-baseline --tune coop_warps=$w > gpurun_out/r02_c2_pairing_w$w.json 2>> gpurun_out/r02_c2_pairing.err
done
for w in 10 14 15; do
python bench.py --workload pairing --steps 5 --warmup 3 --no-cpu-baseline --no-e2e --tune coop_warps=$w > gpurun_out/r02_c2_pairing_w$w.json 2>> gpurun_out/r02_c2_pairing.err
done
python bench.py --workload pairing --steps 5 --warmup 3 --no-cpu-baseline --no-e2e --tune pairing_variant=4 > gpurun_out/r02_c2_pairing_v4.json 2>> gpurun_out/r02_c2_pairing.err
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_coop_pairing -s 3 -c 1 -o gpurun_out/r02_ncu_coop_w12 \
    python bench.py --workload pairing --steps 1 --warmup 3 --no-cpu-baseline --no-e2e --tune coop_warps=12 > /dev/null 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_coop_pairing -s 3 -c 1 -o gpurun_out/r02_ncu_coop_w16 \
    python bench.py --workload pairing --steps 1 --warmup 3 --no-cpu-baseline --no-e2e --tune coop_warps=16 > /dev/null 2>&1
cat gpurun_out/r02_c2_pytest.txt
for f in gpurun_out/r02_c2_pairing_*.json; do echo $f; python -c "
import json,sys
d=json.load(open('$f')); r=d.get('roofline') or {}
print(d['ms_per_step'], d['value'], (d.get('e2e') or {}).get('ms_per_step'), r.get('frac'), r.get('kernel_ms'))"; done
tail -5 gpurun_out/r02_c2_pairing.err
