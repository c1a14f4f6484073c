#!/bin/bash
# 8 GPUs: N = 1, 2, 4, 8 back to back on one box (no parity legs / cpu baseline / config sweep: scaling only)
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-parity --no-configs > gpurun_out/r02_scale_n1.json 2> gpurun_out/r02_scale_n1.err
echo "n1 rc=$?"
for n in 2 4 8; do
timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 2970$n bench.py --gpus $n --steps 10 --warmup 3 --no-cpu-baseline --no-parity --no-configs > gpurun_out/r02_scale_n$n.json 2> gpurun_out/r02_scale_n$n.err
echo "n$n rc=$?"
done
python - <<PY
import json
base=None
for n in (1,2,4,8):
    try:
        d=json.loads(open(f'gpurun_out/r02_scale_n{n}.json').read().strip().splitlines()[-1])
        if n==1: base=d['value']
        print(n, round(d['ms_per_step'],3), round(d['value'],1), 'eff', round(d['value']/(n*base),3), 'e2e', round(d['e2e']['value'],1), d['clocks'])
    except Exception as e: print(n,'ERR',e)
PY
