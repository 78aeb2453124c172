#!/bin/bash
# Round 2, GPU call 16 (1 GPU): cyclotomic squaring with uniform three-term sums + runs of squarings in one call: tests, pairing bench
set -u
mkdir -p gpurun_out
python -m pytest tests -m gpu -q -x -p no:cacheprovider 2>&1 | tail -6 > gpurun_out/r02_c16_pytest.txt
python bench.py --workload pairing --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/r02_c16_pairing.json 2>> gpurun_out/r02_c16.err
for l in 10 13 14; do
python bench.py --workload pairing --log2n $l --steps 5 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/r02_c16_pairing_n$l.json 2>> gpurun_out/r02_c16.err
done
cat gpurun_out/r02_c16_pytest.txt
for f in gpurun_out/r02_c16_*.json; do python - "$f" <<'PY'
import json,sys
d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
e=(d.get('e2e') or {}).get('ms_per_step')
r=d.get('roofline') or {}
print(sys.argv[1], round(d['ms_per_step'],3), '%.4g'%d['value'], 'e2e', e, 'frac', r.get('frac'), {k:round(v,3) for k,v in (r.get('kernel_ms') or {}).items()})
PY
done
tail -n 5 gpurun_out/r02_c16.err
