#!/bin/bash
# final-candidate evidence: full GPU suite, default bench line, launch list of ~2 steps
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -x -q 2>&1 | tail -3 | tee gpurun_out/r02_gpu_tests_32.log
t0=$(date +%s)
timeout 1500 python bench.py > gpurun_out/r02_bench_default.json 2> gpurun_out/r02_bench_default.err
echo "default bench rc=$? wall=$(( $(date +%s) - t0 )) s" | tee gpurun_out/r02_bench_default.wall
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -s 2000 -c 700 --csv --log-file gpurun_out/r02_launches_32.csv python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-parity --no-clocks --no-configs > gpurun_out/r02_launches_32.log 2>&1
echo "ncu rc=$?"; wc -l gpurun_out/r02_launches_32.csv
python - <<PY
import json
d=json.loads(open('gpurun_out/r02_bench_default.json').read().strip().splitlines()[-1])
print(round(d['ms_per_step'],3), round(d['value'],1), 'e2e', round(d['e2e']['value'],1), d['clocks'], d['gpu_launches'])
print(d['roofline']['frac'], d['roofline'].get('frac_net'), d['cpu_baseline']['value'])
print(d['config']['parity']['loss_rel_diff_bf16_vs_fp32'], d['config']['parity']['loss_rel_diff_pdl_on_vs_off'])
PY
