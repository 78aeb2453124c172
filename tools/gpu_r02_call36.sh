#!/bin/bash
# Round 2, GPU call 36 (1 GPU): 2^13..2^15-point G1 MSMs by window width
set -u
mkdir -p gpurun_out
for l in 13 14 15; do for c in 0 8 9 10 11 12 13; do
python bench.py --workload g1_msm --log2n $l --window $c --steps 5 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/r02_c36_g1_n${l}_c$c.json 2>> gpurun_out/r02_c36.err
python - gpurun_out/r02_c36_g1_n${l}_c$c.json <<'PY'
import json,sys
d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
r=d.get('roofline') or {}
print(sys.argv[1], round(d['ms_per_step'],3), {k:round(v,3) for k,v in (r.get('kernel_ms') or {}).items() if 'accum' in k or 'giant' in k or 'horner' in k})
PY
done; done
tail -n 3 gpurun_out/r02_c36.err
