#!/bin/bash
# Round 2, GPU call 30 (1 GPU): compute-sanitizer memcheck + racecheck over the round-2 kernels (six-lane pairing with its
# shared-memory board and group __syncwarp, products, grouped scalar multiplication, cooperative bucket reduction)
set -u
mkdir -p gpurun_out
OUT=gpurun_out/r02_c30_compute_sanitizer.txt
echo "# compute-sanitizer --tool memcheck python -m pytest tests/test_gpu_parity.py -m gpu -k 'pairing_parity or products_shared or mul_batch_items or msm_small or g2_prepared'  (B200, round 2)" > $OUT
timeout 600 compute-sanitizer --tool memcheck python -m pytest tests/test_gpu_parity.py -m gpu -q -p no:cacheprovider -k 'pairing_parity or products_shared or mul_batch_items or msm_small or g2_prepared' 2>&1 | grep -E "COMPUTE-SANITIZER|ERROR SUMMARY|passed|failed|error" | head -20 >> $OUT
echo >> $OUT
echo "# compute-sanitizer --tool racecheck python -m pytest tests/test_gpu_parity.py -m gpu -k 'pairing_parity or products_shared'   (six-lane kernels: board in shared memory, __syncwarp per group)" >> $OUT
timeout 900 compute-sanitizer --tool racecheck python -m pytest tests/test_gpu_parity.py -m gpu -q -p no:cacheprovider -k 'pairing_parity or products_shared' 2>&1 | grep -E "COMPUTE-SANITIZER|RACECHECK SUMMARY|passed|failed|rror" | head -20 >> $OUT
echo >> $OUT
echo "# compute-sanitizer --tool racecheck python -m pytest tests/test_gpu_parity.py -m gpu -k 'msm_small and 1'   (cooperative bucket reduction, shuffles + shared-memory trees)" >> $OUT
timeout 600 compute-sanitizer --tool racecheck python -m pytest tests/test_gpu_parity.py -m gpu -q -p no:cacheprovider -k 'msm_small and 1' 2>&1 | grep -E "COMPUTE-SANITIZER|RACECHECK SUMMARY|passed|failed|rror" | head -20 >> $OUT
cat $OUT
