#!/bin/bash
# 2 GPUs: --check-dp with the Adam-moment gate, N=1 and N=2 bench on the same box; racecheck of smoke() again
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29611 bench.py --gpus 2 --check-dp > gpurun_out/r02_check_dp_n2.json 2> gpurun_out/r02_check_dp_n2.err
echo "check-dp rc=$?"; tail -1 gpurun_out/r02_check_dp_n2.json | cut -c1-1500; tail -3 gpurun_out/r02_check_dp_n2.err
timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-parity --no-configs > gpurun_out/r02_bench_n1_14.json 2> gpurun_out/r02_bench_n1_14.err
echo "n1 rc=$?"
timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29612 bench.py --gpus 2 --steps 10 --warmup 3 --no-parity --no-configs > gpurun_out/r02_bench_n2_14.json 2> gpurun_out/r02_bench_n2_14.err
echo "n2 rc=$?"
python - <<PY
import json
for f in ('n1_14','n2_14'):
    try:
        d=json.loads(open(f'gpurun_out/r02_bench_{f}.json').read().strip().splitlines()[-1])
        print(f, round(d['ms_per_step'],3), round(d['value'],1), round(d['e2e']['value'],1), d['config'].get('loss_last'))
    except Exception as e: print(f,'ERR',e)
PY
timeout 900 compute-sanitizer --tool racecheck --racecheck-report all --print-limit 20 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" > gpurun_out/r02_racecheck_smoke.log 2>&1
echo "racecheck rc=$?"; tail -4 gpurun_out/r02_racecheck_smoke.log
