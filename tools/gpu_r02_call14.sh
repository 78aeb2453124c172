#!/bin/bash
# Round 2, GPU call 14 (1 GPU): products of pairings (shared squaring) vs independent loops; per-shard times of the sharded 2^20 MSM
set -u
mkdir -p gpurun_out
timeout 600 python tools/exp_sweep.py products > gpurun_out/r02_c14_products.jsonl 2> gpurun_out/r02_c14_products.err
for args in "8" "8 g1_glv=2" "8 msm_window=13" "8 msm_window=14" "4" "2"; do
timeout 300 python tools/bench_shard.py $args >> gpurun_out/r02_c14_shards.txt 2>> gpurun_out/r02_c14_shards.err
done
cat gpurun_out/r02_c14_products.jsonl | cut -c1-600
cat gpurun_out/r02_c14_shards.txt
tail -n 3 gpurun_out/r02_c14_products.err gpurun_out/r02_c14_shards.err
