#!/bin/bash
# Round 2, GPU call 12 (1 GPU): shared-memory counting sort (k_msm_sort_sm) on / off; pairing default = six lanes at 2^16, pinned e2e
set -u
mkdir -p gpurun_out
python -m pytest tests -m gpu -q -x -p no:cacheprovider 2>&1 | tail -6 > gpurun_out/r02_c12_pytest.txt
for srt in 1 0; do
python bench.py --workload g1_msm --steps 20 --warmup 3 --no-cpu-baseline --tune msm_sort=$srt > gpurun_out/r02_c12_g1_sort$srt.json 2>> gpurun_out/r02_c12.err
python bench.py --workload g2_msm --steps 6 --warmup 3 --no-cpu-baseline --no-e2e --tune msm_sort=$srt > gpurun_out/r02_c12_g2_sort$srt.json 2>> gpurun_out/r02_c12.err
done
for l in 16 18 22; do
python bench.py --workload g1_msm --log2n $l --steps 10 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/r02_c12_g1_n$l.json 2>> gpurun_out/r02_c12.err
done
python bench.py --workload pairing --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/r02_c12_pairing.json 2>> gpurun_out/r02_c12.err
python bench.py --workload g1_mul --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/r02_c12_g1_mul.json 2>> gpurun_out/r02_c12.err
cat gpurun_out/r02_c12_pytest.txt
for f in gpurun_out/r02_c12_*.json; do python - "$f" <<'PY'
import json,sys
d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
e=(d.get('e2e') or {}).get('ms_per_step')
r=d.get('roofline') or {}
print(sys.argv[1], round(d['ms_per_step'],3), '%.4g'%d['value'], 'e2e', e, 'frac', r.get('frac'), 'exec', r.get('executed_frac'), {k:round(v,3) for k,v in (r.get('kernel_ms') or {}).items()})
PY
done
tail -n 5 gpurun_out/r02_c12.err
