#!/bin/bash
# Round 2, GPU call 4: A/B of the six-lane pairing kernels — called vs inlined Fp2 glue, fused vs split launches
set -u
mkdir -p gpurun_out
run() { # tag
  for sp in 0 1; do for w in 8 12; do
    python bench.py --workload pairing --steps 4 --warmup 3 --no-cpu-baseline --no-e2e --tune coop_warps=$w --tune coop_split=$sp > gpurun_out/r02_c4_$1_s${sp}_w$w.json 2>> gpurun_out/r02_c4.err
  done; done
}
run called
cp bls12_381_b200/csrc/coop12.cuh /tmp/coop12_called.cuh
cp tools/exp/coop12_inline_glue.cuh bls12_381_b200/csrc/coop12.cuh
python -c "from bls12_381_b200 import build; build.build(verbose=True)" >> gpurun_out/r02_c4.err 2>&1
run inline
cp /tmp/coop12_called.cuh bls12_381_b200/csrc/coop12.cuh
for f in gpurun_out/r02_c4_*.json; do python -c "
import json,sys
d=json.load(open('$f')); r=d.get('roofline') or {}
print('$f', round(d['ms_per_step'],2), r.get('kernel_ms'))"; done
tail -3 gpurun_out/r02_c4.err
