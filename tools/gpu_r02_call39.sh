#!/bin/bash
# Round 2, GPU call 39 (1 GPU): ncu --set full of the small-batch pairing path (1024 pairs: k_coop_g2_prepare, Miller loop, final exponentiation, one warp per scheduler)
set -u
mkdir -p gpurun_out
timeout 200 ncu --set full --clock-control none --import-source on -k regex:k_coop --launch-skip 3 --launch-count 3 -f -o gpurun_out/r02_c39_small_pairing python bench.py --workload pairing --log2n 10 --steps 1 --warmup 1 --no-cpu-baseline --no-e2e > gpurun_out/r02_c39.log 2>&1
python tools/ncu_summary.py gpurun_out/r02_c39_small_pairing.ncu-rep > gpurun_out/r02_c39_small_pairing.txt
grep -E "^==|duration|fmaheavy|registers|grid_size|block_size|issue_active|stalled_wait" gpurun_out/r02_c39_small_pairing.txt
