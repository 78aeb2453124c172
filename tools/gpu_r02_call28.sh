#!/bin/bash
# Round 2, GPU call 28 (1 GPU): G2 line coefficients with the lazy-reduction Fp2 build (g2_prepare=1) vs pairing_v4.cu's kernel (0)
set -u
mkdir -p gpurun_out
python -m pytest tests -m gpu -q -x -p no:cacheprovider -k "pairing or prepared" 2>&1 | tail -3
for v in 1 0; do
python bench.py --workload pairing --steps 5 --warmup 3 --no-cpu-baseline --no-e2e --tune g2_prepare=$v --tune coop_chunks=1 > gpurun_out/r02_c28_pairing_prep${v}_ch1.json 2>> gpurun_out/r02_c28.err
python bench.py --workload pairing --steps 5 --warmup 3 --no-cpu-baseline --no-e2e --tune g2_prepare=$v > gpurun_out/r02_c28_pairing_prep${v}_ch3.json 2>> gpurun_out/r02_c28.err
done
for f in gpurun_out/r02_c28_*.json; do python - "$f" <<'PY'
import json,sys
d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
r=d.get('roofline') or {}
print(sys.argv[1], round(d['ms_per_step'],3), '%.4g'%d['value'], 'frac', r.get('frac'), {k:round(v,3) for k,v in (r.get('kernel_ms') or {}).items()})
PY
done
tail -n 3 gpurun_out/r02_c28.err
