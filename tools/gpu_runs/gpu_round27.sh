#!/bin/bash
# programmatic dependent launch: full GPU suite, then the bench with and without it
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
true

for pdl in 1 0; do
WEEDCU_PDL=$pdl timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/bench_pdl$pdl.json 2> gpurun_out/bench_pdl$pdl.err
echo "bench pdl=$pdl rc=$?"; python - <<PY
import json
d=json.loads(open('gpurun_out/bench_pdl$pdl.json').read().strip().splitlines()[-1])
print(d['ms_per_step'], d['value'], d['e2e']['value'], d['config']['loss_first'], d['config']['loss_last'])
print({k: round(v['ms_per_step'],3) for k,v in d['kernel_breakdown'].items()})
print(d['clocks'], d['host'])
PY
done
