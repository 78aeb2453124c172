#!/bin/bash
# Round 2, GPU call 9 (1 GPU): lane-cooperative bucket reduction + H2D overlap — tests, G1/G2 MSM timing both reductions
set -u
mkdir -p gpurun_out
python -m pytest tests -m gpu -q -x -p no:cacheprovider 2>&1 | tail -6 > gpurun_out/r02_c9_pytest.txt
for wl in g1_msm g2_msm; do for r in 1 0; do
python bench.py --workload $wl --steps 10 --warmup 3 --no-cpu-baseline --tune msm_reduce=$r > gpurun_out/r02_c9_${wl}_red$r.json 2>> gpurun_out/r02_c9.err
done; done
for l in 16 18 22; do
python bench.py --workload g1_msm --log2n $l --steps 10 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/r02_c9_g1_n$l.json 2>> gpurun_out/r02_c9.err
done
cat gpurun_out/r02_c9_pytest.txt
for f in gpurun_out/r02_c9_g*.json; do python - "$f" <<'PY'
import json,sys
d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
e=(d.get('e2e') or {}).get('ms_per_step')
print(sys.argv[1], round(d['ms_per_step'],3), round(d['value']), 'e2e', e, {k:round(v,3) for k,v in (d['roofline'] or {}).get('kernel_ms',{}).items() if 'reduce' in k or 'fold' in k or 'horner' in k or 'accum' in k})
PY
done
tail -n 3 gpurun_out/r02_c9.err
