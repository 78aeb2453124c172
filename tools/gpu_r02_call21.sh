#!/bin/bash
# Round 2, GPU call 21 (1 GPU): ncu --set full of the dominant kernels in their final round-2 form
set -u
mkdir -p gpurun_out
ncu --set full --clock-control none --import-source on -k regex:k_coop_pairing --launch-skip 2 --launch-count 2 -f -o gpurun_out/r02_c21_coop_pairing python bench.py --workload pairing --steps 1 --warmup 1 --no-cpu-baseline --no-e2e > gpurun_out/r02_c21_p.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:k_msm_accumulate_g2sm --launch-skip 4 --launch-count 1 -f -o gpurun_out/r02_c21_g2_acc python bench.py --workload g2_msm --steps 1 --warmup 1 --no-cpu-baseline --no-e2e > gpurun_out/r02_c21_g2.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:k_msm_accumulate --launch-skip 4 --launch-count 1 -f -o gpurun_out/r02_c21_g1_acc python bench.py --workload g1_msm --steps 1 --warmup 1 --no-cpu-baseline --no-e2e > gpurun_out/r02_c21_g1.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:k_mul_batch_grp --launch-skip 1 --launch-count 1 -f -o gpurun_out/r02_c21_mul_grp python bench.py --workload g1_mul --steps 2 --warmup 1 --no-cpu-baseline --no-e2e > gpurun_out/r02_c21_mul.log 2>&1
ls -la gpurun_out/*.ncu-rep
for f in coop_pairing g2_acc g1_acc mul_grp; do python tools/ncu_summary.py gpurun_out/r02_c21_$f.ncu-rep > gpurun_out/r02_c21_$f.txt; done
grep -E "^==|duration|fmaheavy|dram__bytes|registers" gpurun_out/r02_c21_*.txt
