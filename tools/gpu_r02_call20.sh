#!/bin/bash
# Round 2, GPU call 20 (1 GPU): main line of the MSM on a highest-priority stream (msm_priority) on / off
set -u
mkdir -p gpurun_out
for p in 1 0; do
python bench.py --workload g2_msm --steps 6 --warmup 3 --no-cpu-baseline --tune msm_priority=$p > gpurun_out/r02_c20_g2_pri$p.json 2>> gpurun_out/r02_c20.err
python bench.py --workload g1_msm --steps 20 --warmup 3 --no-cpu-baseline --tune msm_priority=$p > gpurun_out/r02_c20_g1_pri$p.json 2>> gpurun_out/r02_c20.err
done
python -m pytest tests -m gpu -q -x -p no:cacheprovider -k "msm" 2>&1 | tail -3
for f in gpurun_out/r02_c20_*.json; do python - "$f" <<'PY'
import json,sys
d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
r=d.get('roofline') or {}
print(sys.argv[1], round(d['ms_per_step'],3), '%.4g'%d['value'], 'e2e', (d.get('e2e') or {}).get('ms_per_step'), {k:round(v,3) for k,v in (r.get('kernel_ms') or {}).items() if 'reduce' in k or 'fold' in k or 'horner' in k or 'accum' in k})
PY
done
tail -n 3 gpurun_out/r02_c20.err
