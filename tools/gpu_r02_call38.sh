#!/bin/bash
# Round 2, GPU call 38 (2 GPUs): the split-sum property check of bench.py (what config 5 uses at 2^24), forced at 2^18, both sharding modes
set -u
mkdir -p gpurun_out
for m in points window; do
B200_BENCH_FORCE_SPLIT_CHECK=1 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29551 bench.py --gpus 2 --workload g1_msm --log2n 18 --steps 5 --warmup 3 --shard $m --no-cpu-baseline > gpurun_out/r02_c38_n2_$m.json 2> gpurun_out/r02_c38_n2_$m.err; echo "rc=$?" >> gpurun_out/r02_c38_n2_$m.err
python - gpurun_out/r02_c38_n2_$m.json <<'PY'
import json,sys
d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
print(sys.argv[1], round(d['ms_per_step'],3), 'e2e', (d.get('e2e') or {}).get('ms_per_step'), d['config'].get('sharded_result_checked'), 'split:', d['config'].get('split_sum_checked'))
PY
tail -n 2 gpurun_out/r02_c38_n2_$m.err
done
B200_BENCH_FORCE_SPLIT_CHECK=1 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29552 bench.py --gpus 2 --workload g2_msm --log2n 16 --steps 3 --warmup 3 --no-cpu-baseline --no-e2e 2> gpurun_out/r02_c38_g2.err | python -c "import json,sys; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('g2', round(d['ms_per_step'],3), d['config'].get('sharded_result_checked'), 'split:', d['config'].get('split_sum_checked'))"
