#!/bin/bash
# Round 2, GPU call 15 (1 GPU): per-shard times of the sharded 2^20 G1 MSM (what each rank of an N-GPU run executes)
set -u
mkdir -p gpurun_out
rm -f gpurun_out/r02_c15_shards.txt
for args in "8" "8 g1_glv=2" "8 msm_window=13" "8 msm_window=14" "8 msm_tail_groups=0" "4" "2"; do
timeout 300 python tools/bench_shard.py $args >> gpurun_out/r02_c15_shards.txt 2>> gpurun_out/r02_c15_shards.err
done
cat gpurun_out/r02_c15_shards.txt
tail -n 3 gpurun_out/r02_c15_shards.err
