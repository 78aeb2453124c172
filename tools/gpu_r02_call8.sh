#!/bin/bash
# Round 2, GPU call 8 (8 GPUs): strong scaling of the 2^20 G1 MSM inside the library — sharding mode x window sweep; default line
set -u
mkdir -p gpurun_out
run() { # tag args...
  tag=$1; shift
  python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29520 bench.py --gpus 8 --steps 10 --warmup 3 --workload g1_msm --no-e2e "$@" > gpurun_out/r02_c8_$tag.json 2> gpurun_out/r02_c8_$tag.err
}
run points_auto --shard points
run points_c12 --shard points --window 12
run points_c14 --shard points --window 14
run points_c16 --shard points --window 16
run window_c16 --shard window
python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29521 bench.py --gpus 8 --steps 10 --warmup 3 > gpurun_out/r02_c8_all_n8.json 2> gpurun_out/r02_c8_all_n8.err; echo "rc=$?" >> gpurun_out/r02_c8_all_n8.err
for f in gpurun_out/r02_c8_*.json; do python - "$f" <<'PY'
import json,sys
try:
    d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    e=(d.get('e2e') or {}).get('ms_per_step')
    print(sys.argv[1], round(d['ms_per_step'],3), round(d['value']), 'e2e', e, {k:round(v,3) for k,v in (d['roofline'] or {}).get('kernel_ms',{}).items()})
    for k,c in d.get('configs',{}).items(): print('  ',k, round(c['ms_per_step'],3), round(c['value']), 'e2e', (c['e2e'] or {}).get('ms_per_step'))
except Exception as e: print(sys.argv[1], 'ERR', e)
PY
done
tail -n 3 gpurun_out/r02_c8_all_n8.err
