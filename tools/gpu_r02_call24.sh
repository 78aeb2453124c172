#!/bin/bash
# Round 2, GPU call 24 (2 GPUs): the N > 1 path after this round's changes — tests on rank-visible GPUs, 2-GPU bench line (all configs)
set -u
mkdir -p gpurun_out
python -m pytest tests -m gpu -q -x -p no:cacheprovider 2>&1 | tail -4 > gpurun_out/r02_c24_pytest.txt
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r02_c24_smoke.txt 2>&1
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus 2 --steps 10 --warmup 3 > gpurun_out/r02_c24_bench_n2.json 2> gpurun_out/r02_c24_bench_n2.err; echo "rc=$?" >> gpurun_out/r02_c24_bench_n2.err
cat gpurun_out/r02_c24_pytest.txt; tail -2 gpurun_out/r02_c24_smoke.txt
python - <<'PY'
import json
d=json.loads(open('gpurun_out/r02_c24_bench_n2.json').read().strip().splitlines()[-1])
def show(name,c):
    r=c.get('roofline') or {}
    print(name,'ms',round(c['ms_per_step'],3),'value','%.4g'%c['value'],'e2e',round(c['e2e']['ms_per_step'],3) if c.get('e2e') else None, c['config'].get('sharding'), c['config'].get('sharded_result_checked'))
show('g1_msm',d)
for k,c in d.get('configs',{}).items(): show(k,c)
PY
tail -n 4 gpurun_out/r02_c24_bench_n2.err
