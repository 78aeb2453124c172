#!/bin/bash
# Round 2, GPU call 19 (1 GPU): bucket-reduction mode (0 thread per chunk, 1 cooperative for the exposed last group, 2 cooperative everywhere) for G2 and G1
set -u
mkdir -p gpurun_out
for r in 1 2; do
python bench.py --workload g2_msm --steps 6 --warmup 3 --no-cpu-baseline --no-e2e --tune msm_reduce=$r > gpurun_out/r02_c19_g2_red$r.json 2>> gpurun_out/r02_c19.err
python bench.py --workload g1_msm --steps 10 --warmup 3 --no-cpu-baseline --no-e2e --tune msm_reduce=$r > gpurun_out/r02_c19_g1_red$r.json 2>> gpurun_out/r02_c19.err
done
for f in gpurun_out/r02_c19_*.json; do python - "$f" <<'PY'
import json,sys
d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
r=d.get('roofline') or {}
print(sys.argv[1], round(d['ms_per_step'],3), '%.4g'%d['value'], {k:round(v,3) for k,v in (r.get('kernel_ms') or {}).items() if 'reduce' in k or 'fold' in k or 'horner' in k or 'accum' in k})
PY
done
tail -n 3 gpurun_out/r02_c19.err
