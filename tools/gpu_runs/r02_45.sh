#!/bin/bash
# final verification: full GPU suite, smoke(), default bench line
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -x -q 2>&1 | tail -3 | tee gpurun_out/r02_gpu_tests_45.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
t0=$(date +%s)
timeout 1500 python bench.py > gpurun_out/r02_bench_default.json 2> gpurun_out/r02_bench_default.err
echo "default bench rc=$? wall=$(( $(date +%s) - t0 )) s" | tee gpurun_out/r02_bench_default.wall
python - <<PY
import json
d=json.loads(open('gpurun_out/r02_bench_default.json').read().strip().splitlines()[-1])
print(round(d['ms_per_step'],3), round(d['value'],1), 'e2e', round(d['e2e']['value'],1), d['clocks'], d['gpu_launches'])
print(d['roofline']['frac'], d['roofline'].get('frac_net'), d['roofline']['traffic'], d['cpu_baseline']['value'])
print(d['config']['parity']['loss_rel_diff_bf16_vs_fp32'], d['config']['parity']['within_bf16_bound'], d['config']['parity']['loss_rel_diff_pdl_on_vs_off'])
ex=d['extra']['configs']; print({k:(v.get('ms_per_step')) for k,v in ex.items() if k in ('c2','c4')}, ex['decode']['tokens_per_s'])
PY
