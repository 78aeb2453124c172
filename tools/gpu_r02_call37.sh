#!/bin/bash
# Round 2, GPU call 37 (1 GPU): MSM window heuristic that models the light last buckets of the top window — tests + small sizes
set -u
mkdir -p gpurun_out
python -m pytest tests -m gpu -q -x -p no:cacheprovider -k "msm or multi or smoke" 2>&1 | tail -3
rm -f gpurun_out/r02_c37_sizes.txt
for l in 10 12 13 14 15 16 17 18 20; do
python bench.py --workload g1_msm --log2n $l --steps 5 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/r02_c37_g1_n$l.json 2>> gpurun_out/r02_c37.err
python - gpurun_out/r02_c37_g1_n$l.json <<'PY' | tee -a gpurun_out/r02_c37_sizes.txt
import json,sys
d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
r=d.get('roofline') or {}
print(sys.argv[1], round(d['ms_per_step'],3), '%.4g'%d['value'], {k:round(v,3) for k,v in (r.get('kernel_ms') or {}).items() if 'accum' in k or 'horner' in k})
PY
done
python bench.py --workload g2_msm --log2n 14 --steps 5 --warmup 3 --no-cpu-baseline --no-e2e 2>> gpurun_out/r02_c37.err | python -c "import json,sys; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('g2 2^14', round(d['ms_per_step'],3))"
tail -n 3 gpurun_out/r02_c37.err
