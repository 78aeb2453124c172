#!/bin/bash
# First GPU call of the next round (run under gpurun from the repo root, ~20-25 GPU-minutes):
#   /usr/local/graft/bin/gpurun --timeout 2400 -- 'bash tools/gpu_first_call_round2.sh'
# 1. the validated gate, 2. the two rows written after round 1's GPU budget ran out (scalar field / NTT, hash to curve:
# expect XPASS — then delete the xfail markers in tests/test_gpu_zz_*.py), 3. their first measurements and ncu captures.
set -u
mkdir -p gpurun_out
python -m pytest tests -m gpu -q -rxXf -p no:cacheprovider 2>&1 | tail -40 > gpurun_out/r02_first_pytest.txt
python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" > gpurun_out/r02_first_smoke.txt 2>&1
python tools/bench_fr_ntt.py --log-n 24 --steps 10 --warmup 3 > gpurun_out/r02_fr_ntt_2p24.json 2> gpurun_out/r02_fr_ntt_2p24.err
python tools/bench_fr_ntt.py --log-n 20 --steps 20 --warmup 3 > gpurun_out/r02_fr_ntt_2p20.json 2>> gpurun_out/r02_fr_ntt_2p24.err
python bench.py --steps 10 --warmup 3 > gpurun_out/r02_bench_g1msm.json 2> gpurun_out/r02_bench_g1msm.err
# the experimental dual-stream pairing kernels against the default ones (same workload, same process conditions)
python bench.py --workload pairing --steps 5 --warmup 3 > gpurun_out/r02_bench_pairing_v4.json 2> gpurun_out/r02_bench_pairing.err
python bench.py --workload pairing --steps 5 --warmup 3 --tune pairing_variant=5 > gpurun_out/r02_bench_pairing_v5.json 2>> gpurun_out/r02_bench_pairing.err
python bench.py --workload pairing --steps 5 --warmup 3 --tune pairing_variant=6 > gpurun_out/r02_bench_pairing_v6.json 2>> gpurun_out/r02_bench_pairing.err
# default G2 MSM vs the experimental second build of the MSM unit (row-alternated lazy Fp2 multiply)
python tools/bench_g2_msm_variant.py --log-n 20 --steps 5 --warmup 2 > gpurun_out/r02_g2_msm_variant.json 2> gpurun_out/r02_g2_msm_variant.err
# launch list + one full capture of the NTT pass kernel (skip the table-building launches of the first call)
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file gpurun_out/r02_ncu_launches_fr_ntt.csv \
    python tools/bench_fr_ntt.py --log-n 22 --steps 2 --warmup 1 --cpu-log-n 12 > /dev/null 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_fr_ntt_pass -s 8 -c 1 -o gpurun_out/r02_ncu_fr_ntt_pass \
    python tools/bench_fr_ntt.py --log-n 22 --steps 1 --warmup 1 --cpu-log-n 12 > /dev/null 2>&1
# per-kernel times of the pairing variants (ncu serialises and runs cold: compare shares, not absolutes)
for v in 4 5 6; do
  timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file gpurun_out/r02_ncu_launches_pairing_v$v.csv \
      python bench.py --workload pairing --steps 1 --warmup 1 --tune pairing_variant=$v > /dev/null 2>&1
done
cat gpurun_out/r02_first_pytest.txt | tail -15
cat gpurun_out/r02_fr_ntt_2p24.json
