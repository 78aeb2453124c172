#!/bin/bash
# Round 2, GPU call 27 (1 GPU): window shards with the 4-blocks-per-SM bucket kernel for launches of 1.0 .. 1.33 waves
set -u
mkdir -p gpurun_out
rm -f gpurun_out/r02_c27_shards.txt
python -m pytest tests -m gpu -q -x -p no:cacheprovider -k "msm or multi" 2>&1 | tail -3
for args in "8" "8 g1_one_wave=0" "4" "2"; do
timeout 300 python tools/bench_shard.py $args >> gpurun_out/r02_c27_shards.txt 2>> gpurun_out/r02_c27_shards.err
done
cat gpurun_out/r02_c27_shards.txt
tail -n 3 gpurun_out/r02_c27_shards.err
