#!/bin/bash
# Round 2, GPU call 5: phase-specific six-lane kernels; small-batch comparison against the one-thread-per-pairing kernels
set -u
mkdir -p gpurun_out
python -m pytest tests -m gpu -q -x -p no:cacheprovider 2>&1 | tail -5 > gpurun_out/r02_c5_pytest.txt
for w in 8 12; do
python bench.py --workload pairing --steps 4 --warmup 3 --no-cpu-baseline --no-e2e --tune coop_warps=$w > gpurun_out/r02_c5_pairing_w$w.json 2>> gpurun_out/r02_c5.err
done
python bench.py --workload pairing --steps 4 --warmup 3 --no-cpu-baseline --no-e2e --tune coop_split=0 > gpurun_out/r02_c5_pairing_fused.json 2>> gpurun_out/r02_c5.err
for l in 10 12 13 14 15; do
python bench.py --workload pairing --log2n $l --steps 5 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/r02_c5_pairing_n${l}_v7.json 2>> gpurun_out/r02_c5.err
python bench.py --workload pairing --log2n $l --steps 5 --warmup 3 --no-cpu-baseline --no-e2e --tune pairing_variant=4 > gpurun_out/r02_c5_pairing_n${l}_v4.json 2>> gpurun_out/r02_c5.err
done
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_coop_pairing -s 6 -c 2 -o gpurun_out/r02_ncu_coop3_w12 \
    python bench.py --workload pairing --steps 1 --warmup 3 --no-cpu-baseline --no-e2e > /dev/null 2>&1
cat gpurun_out/r02_c5_pytest.txt
for f in gpurun_out/r02_c5_pairing_*.json; do python -c "
import json,sys
d=json.load(open('$f')); r=d.get('roofline') or {}
print('$f', round(d['ms_per_step'],3), round(d['value']), r.get('kernel_ms'))"; done
tail -3 gpurun_out/r02_c5.err
