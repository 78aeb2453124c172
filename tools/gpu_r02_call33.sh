#!/bin/bash
# Round 2, GPU call 33 (1 GPU): small MSMs — buckets per thread of the overlapped reduction (partials the Horner kernel sums serially)
set -u
mkdir -p gpurun_out
python -m pytest tests -m gpu -q -x -p no:cacheprovider -k "msm" 2>&1 | tail -3
for l in 14 16 17 18 20; do for m in 1 8 16; do
python bench.py --workload g1_msm --log2n $l --steps 10 --warmup 3 --no-cpu-baseline --no-e2e --tune msm_reduce_min_chunk=$m > gpurun_out/r02_c33_g1_n${l}_mc$m.json 2>> gpurun_out/r02_c33.err
done; done
for f in gpurun_out/r02_c33_*.json; do python - "$f" <<'PY'
import json,sys
d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
r=d.get('roofline') or {}
print(sys.argv[1], round(d['ms_per_step'],3), '%.4g'%d['value'], {k:round(v,3) for k,v in (r.get('kernel_ms') or {}).items() if 'reduce' in k or 'horner' in k or 'accum' in k})
PY
done
tail -n 3 gpurun_out/r02_c33.err
