#!/bin/bash
# Round 2, GPU call 6 (2 GPUs): multi-GPU inside the library — tests, torchrun check, bench at N = 2 in both sharding modes
set -u
mkdir -p gpurun_out
nvidia-smi -L > gpurun_out/r02_c6_gpus.txt
python -m pytest tests -m gpu -q -x -p no:cacheprovider 2>&1 | tail -8 > gpurun_out/r02_c6_pytest.txt
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 tools/multi_gpu_check.py > gpurun_out/r02_c6_check.txt 2>&1; echo "check rc=$?" >> gpurun_out/r02_c6_check.txt
for sh in points window; do
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 2 --steps 10 --warmup 3 --workload g1_msm --shard $sh > gpurun_out/r02_c6_g1msm_n2_$sh.json 2> gpurun_out/r02_c6_g1msm_n2_$sh.err; echo "rc=$?" >> gpurun_out/r02_c6_g1msm_n2_$sh.err
done
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29513 bench.py --gpus 2 --steps 5 --warmup 3 > gpurun_out/r02_c6_all_n2.json 2> gpurun_out/r02_c6_all_n2.err; echo "rc=$?" >> gpurun_out/r02_c6_all_n2.err
cat gpurun_out/r02_c6_pytest.txt; tail -6 gpurun_out/r02_c6_check.txt
for f in gpurun_out/r02_c6_g1msm_n2_points.json gpurun_out/r02_c6_g1msm_n2_window.json gpurun_out/r02_c6_all_n2.json; do python - "$f" <<'PY'
import json,sys
try:
    d=json.load(open(sys.argv[1]))
    print(sys.argv[1], round(d['ms_per_step'],3), round(d['value']), 'e2e', round(d['e2e']['ms_per_step'],3), d['config']['sharding'][:30], (d['roofline'] or {}).get('kernel_ms'))
    for k,c in d.get('configs',{}).items(): print('  ',k, round(c['ms_per_step'],3), round(c['value']), (c['e2e'] or {}).get('ms_per_step'))
except Exception as e: print(sys.argv[1], 'ERR', e)
PY
done
tail -4 gpurun_out/r02_c6_g1msm_n2_points.err gpurun_out/r02_c6_all_n2.err
