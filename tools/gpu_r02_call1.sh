#!/bin/bash
# Round 2, GPU call 1: gate + first measurements of everything written after round 1's GPU budget ran out.
set -u
mkdir -p gpurun_out
python -m pytest tests -m gpu -q -rxXf -p no:cacheprovider 2>&1 | tail -40 > gpurun_out/r02_first_pytest.txt
python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" > gpurun_out/r02_first_smoke.txt 2>&1
python tools/bench_fr_ntt.py --log-n 24 --steps 10 --warmup 3 > gpurun_out/r02_fr_ntt_2p24.json 2> gpurun_out/r02_fr_ntt_2p24.err
python tools/bench_fr_ntt.py --log-n 20 --steps 20 --warmup 3 > gpurun_out/r02_fr_ntt_2p20.json 2>> gpurun_out/r02_fr_ntt_2p24.err
python bench.py --steps 10 --warmup 3 > gpurun_out/r02_bench_g1msm.json 2> gpurun_out/r02_bench_g1msm.err
python bench.py --workload pairing --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/r02_bench_pairing_v4.json 2> gpurun_out/r02_bench_pairing.err
python bench.py --workload pairing --steps 5 --warmup 3 --no-cpu-baseline --tune pairing_variant=5 > gpurun_out/r02_bench_pairing_v5.json 2>> gpurun_out/r02_bench_pairing.err
python bench.py --workload pairing --steps 5 --warmup 3 --no-cpu-baseline --tune pairing_variant=6 > gpurun_out/r02_bench_pairing_v6.json 2>> gpurun_out/r02_bench_pairing.err
python tools/bench_g2_msm_variant.py --log-n 20 --steps 5 --warmup 2 > gpurun_out/r02_g2_msm_variant.json 2> gpurun_out/r02_g2_msm_variant.err
python bench.py --workload g2_msm --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/r02_bench_g2msm.json 2> gpurun_out/r02_bench_g2msm.err
python bench.py --workload g1_mul --steps 20 --warmup 3 --no-cpu-baseline > gpurun_out/r02_bench_g1mul.json 2> gpurun_out/r02_bench_g1mul.err
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file gpurun_out/r02_ncu_launches_fr_ntt.csv \
    python tools/bench_fr_ntt.py --log-n 22 --steps 2 --warmup 1 --cpu-log-n 12 > /dev/null 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_fr_ntt_pass -s 8 -c 1 -o gpurun_out/r02_ncu_fr_ntt_pass \
    python tools/bench_fr_ntt.py --log-n 22 --steps 1 --warmup 1 --cpu-log-n 12 > /dev/null 2>&1
# the SHIPPED G2 bucket kernel (k_msm_accumulate_g2sm<2>): first launch of a 2^20 MSM (6-window group)
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_msm_accumulate_g2sm -s 3 -c 1 -o gpurun_out/r02_ncu_g2sm \
    python bench.py --workload g2_msm --steps 1 --warmup 3 --no-cpu-baseline --no-e2e > /dev/null 2>&1
tail -15 gpurun_out/r02_first_pytest.txt
cat gpurun_out/r02_fr_ntt_2p24.json gpurun_out/r02_bench_pairing_v?.json gpurun_out/r02_g2_msm_variant.json | cut -c1-600
