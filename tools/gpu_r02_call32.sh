#!/bin/bash
# Round 2, GPU call 32 (1 GPU): small pairing batches spread over all SMs (as few warps per scheduler as possible) + six-lane prepare
set -u
mkdir -p gpurun_out
python -m pytest tests -m gpu -q -x -p no:cacheprovider -k "pairing or prepared or products or gt" 2>&1 | tail -3
for l in 8 10 11 12 13 14 16; do
python bench.py --workload pairing --log2n $l --steps 5 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/r02_c32_pairing_n${l}.json 2>> gpurun_out/r02_c32.err
done
timeout 300 python tools/exp_sweep.py products --reps 3 > gpurun_out/r02_c32_products.jsonl 2>> gpurun_out/r02_c32.err
for f in gpurun_out/r02_c32_*.json; do python - "$f" <<'PY'
import json,sys
d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
r=d.get('roofline') or {}
print(sys.argv[1], round(d['ms_per_step'],3), '%.4g'%d['value'], {k:round(v,3) for k,v in (r.get('kernel_ms') or {}).items()})
PY
done
cut -c1-330 gpurun_out/r02_c32_products.jsonl
tail -n 3 gpurun_out/r02_c32.err
