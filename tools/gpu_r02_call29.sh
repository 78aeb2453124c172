#!/bin/bash
# Round 2, GPU call 29 (8 GPUs): the bench line the driver's SCALE step runs at N = 8 (all configs incl. config 5 = 2^24)
set -u
mkdir -p gpurun_out
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29541 bench.py --gpus 8 --steps 10 --warmup 3 > gpurun_out/r02_c29_bench_n8.json 2> gpurun_out/r02_c29_bench_n8.err; echo "rc=$?" >> gpurun_out/r02_c29_bench_n8.err
python - <<'PY'
import json
d=json.loads(open('gpurun_out/r02_c29_bench_n8.json').read().strip().splitlines()[-1])
def show(name,c):
    print(name,'ms',round(c['ms_per_step'],3),'value','%.4g'%c['value'],'e2e',round(c['e2e']['ms_per_step'],3) if c.get('e2e') else None, c['config'].get('sharding','')[:40], c['config'].get('sharded_result_checked'))
show('g1_msm',d)
for k,c in d.get('configs',{}).items(): show(k,c)
PY
tail -n 4 gpurun_out/r02_c29_bench_n8.err
