#!/bin/bash
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/r01_bench_n1_v17.json 2>/dev/null
python - <<PY
import json
d=json.loads(open('gpurun_out/r01_bench_n1_v17.json').read().strip().splitlines()[-1]); print('n1', round(d['ms_per_step'],3), round(d['value'],1))
PY
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29531 bench.py --gpus 2 --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/r01_bench_n2_v17.json 2> gpurun_out/n2.err
python - <<PY
import json
d=json.loads(open('gpurun_out/r01_bench_n2_v17.json').read().strip().splitlines()[-1]); print('n2', round(d['ms_per_step'],3), round(d['value'],1), d['config']['loss_last'])
PY
