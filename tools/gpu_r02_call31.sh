#!/bin/bash
# Round 2, GPU call 31 (1 GPU): G2 line coefficients with six lanes per Q (k_coop_g2_prepare) vs one thread per Q, by batch size
set -u
mkdir -p gpurun_out
python -m pytest tests -m gpu -q -x -p no:cacheprovider -k "pairing or prepared or products" 2>&1 | tail -3
for l in 10 12 13 14 15; do for m in 0 1000000; do
python bench.py --workload pairing --log2n $l --steps 5 --warmup 3 --no-cpu-baseline --no-e2e --tune coop_prepare_max=$m > gpurun_out/r02_c31_pairing_n${l}_prep$m.json 2>> gpurun_out/r02_c31.err
done; done
for f in gpurun_out/r02_c31_*.json; do python - "$f" <<'PY'
import json,sys
d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
r=d.get('roofline') or {}
print(sys.argv[1], round(d['ms_per_step'],3), '%.4g'%d['value'], {k:round(v,3) for k,v in (r.get('kernel_ms') or {}).items()})
PY
done
tail -n 3 gpurun_out/r02_c31.err
