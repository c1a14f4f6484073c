#!/bin/bash
# 2 GPUs: data-parallel equality (--check-dp) + N=1 / N=2 bench on the same box (Adam chained per bucket vs not)
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
run2() { python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port $1 bench.py --gpus 2 "${@:2}"; }
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29611 bench.py --gpus 2 --check-dp > gpurun_out/r02_check_dp_n2.json 2> gpurun_out/r02_check_dp_n2.err
echo "check-dp rc=$?"; cat gpurun_out/r02_check_dp_n2.json | tail -1 | cut -c1-900; tail -3 gpurun_out/r02_check_dp_n2.err
timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-parity > gpurun_out/r02_bench_n1_09.json 2> gpurun_out/r02_bench_n1_09.err
echo "n1 rc=$?"
timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29612 bench.py --gpus 2 --steps 10 --warmup 3 > gpurun_out/r02_bench_n2_09.json 2> gpurun_out/r02_bench_n2_09.err
echo "n2 rc=$?"
WH_DP_CHAIN_ADAM=0 timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29613 bench.py --gpus 2 --steps 10 --warmup 3 > gpurun_out/r02_bench_n2_nochain_09.json 2> gpurun_out/r02_bench_n2_nochain_09.err
echo "n2 nochain rc=$?"
python - <<PY
import json
for f in ('n1_09','n2_09','n2_nochain_09'):
    try:
        d=json.loads(open(f'gpurun_out/r02_bench_{f}.json').read().strip().splitlines()[-1])
        print(f, round(d['ms_per_step'],3), round(d['value'],1), round(d['e2e']['value'],1), d['config']['loss_last'])
    except Exception as e: print(f,'ERR',e)
PY
