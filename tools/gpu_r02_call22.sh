#!/bin/bash
# Round 2, GPU call 22 (1 GPU): six-lane pairing batch as chunks on two streams (coop_chunks sweep)
set -u
mkdir -p gpurun_out
python -m pytest tests -m gpu -q -x -p no:cacheprovider -k "pairing" 2>&1 | tail -3
for c in 1 2 3 4 6 8; do
python bench.py --workload pairing --steps 5 --warmup 3 --no-cpu-baseline --no-e2e --tune coop_chunks=$c > gpurun_out/r02_c22_pairing_ch$c.json 2>> gpurun_out/r02_c22.err
done
for f in gpurun_out/r02_c22_*.json; do python - "$f" <<'PY'
import json,sys
d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
r=d.get('roofline') or {}
print(sys.argv[1], round(d['ms_per_step'],3), '%.4g'%d['value'], 'frac', r.get('frac'), {k:round(v,3) for k,v in (r.get('kernel_ms') or {}).items()})
PY
done
tail -n 3 gpurun_out/r02_c22.err
