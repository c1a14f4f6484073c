#!/bin/bash
# round 2, call 3: full GPU suite with the epilogue fusions + bench (breakdown) with fusions on and off
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
timeout 900 python -m pytest tests -q -m gpu -x > gpurun_out/r02_gpu_tests_03.log 2>&1
echo "gpu tests rc=$?"; tail -15 gpurun_out/r02_gpu_tests_03.log
timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-parity > gpurun_out/r02_bench_03_on.json 2> gpurun_out/r02_bench_03_on.err
echo "bench on rc=$?"; tail -3 gpurun_out/r02_bench_03_on.err
WEED_B200_EPILOGUE=0 timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-parity > gpurun_out/r02_bench_03_off.json 2> gpurun_out/r02_bench_03_off.err
echo "bench off rc=$?"; tail -3 gpurun_out/r02_bench_03_off.err
python - <<PY
import json
for f in ('on','off'):
    try:
        d=json.loads(open(f'gpurun_out/r02_bench_03_{f}.json').read().strip().splitlines()[-1])
        print(f, round(d['ms_per_step'],3), round(d['value'],1), round(d['e2e']['value'],1), d['config']['loss_first'], d['config']['loss_last'], d['gpu_launches'])
        print({k:(round(v['ms_per_step'],3), v['launches_per_step']) for k,v in d['kernel_breakdown'].items()})
    except Exception as e: print(f, 'ERR', e)
PY
