#!/bin/bash
# Round 2, GPU call 25 (1 GPU): window shards — one window per group + cooperative reduction for every group (msm_coop_nloc)
set -u
mkdir -p gpurun_out
rm -f gpurun_out/r02_c25_shards.txt
for args in "8 msm_tail_groups=1 msm_coop_nloc=3" "8 msm_tail_groups=0 msm_coop_nloc=3" "8" "4 msm_coop_nloc=4" "4 msm_coop_nloc=4 msm_tail_groups=1" "4" "2 msm_coop_nloc=8" "2"; do
timeout 300 python tools/bench_shard.py $args >> gpurun_out/r02_c25_shards.txt 2>> gpurun_out/r02_c25_shards.err
done
cat gpurun_out/r02_c25_shards.txt
tail -n 3 gpurun_out/r02_c25_shards.err
