#!/bin/bash
# Round 2, GPU call 10 (1 GPU): fused-shift REDC + grouped scalar-mul kernel: tests; start-up stagger sweep of the pairing kernels;
# scalar-mul shape sweep; G2 MSM with the new REDC
set -u
mkdir -p gpurun_out
python -m pytest tests -m gpu -q -x -p no:cacheprovider 2>&1 | tail -6 > gpurun_out/r02_c10_pytest.txt
timeout 300 python tools/exp_sweep.py pairing --warps 12,8 > gpurun_out/r02_c10_pairing.jsonl 2> gpurun_out/r02_c10_pairing.err
timeout 300 python tools/exp_sweep.py mul > gpurun_out/r02_c10_mul.jsonl 2> gpurun_out/r02_c10_mul.err
timeout 300 python bench.py --workload g2_msm --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/r02_c10_g2_msm.json 2> gpurun_out/r02_c10_g2.err
cat gpurun_out/r02_c10_pytest.txt
cat gpurun_out/r02_c10_pairing.jsonl | cut -c1-400
cat gpurun_out/r02_c10_mul.jsonl
tail -n 3 gpurun_out/r02_c10_pairing.err gpurun_out/r02_c10_mul.err gpurun_out/r02_c10_g2.err
python - <<'PY'
import json
d=json.loads(open('gpurun_out/r02_c10_g2_msm.json').read().strip().splitlines()[-1])
print('g2_msm', round(d['ms_per_step'],3), (d.get('e2e') or {}).get('ms_per_step'), {k:round(v,3) for k,v in d['roofline']['kernel_ms'].items()})
PY
