#!/bin/bash
# round-2 evidence: default bench line (parity legs, cpu baseline, extra.configs) with its wall time
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
t0=$(date +%s)
timeout 1500 python bench.py > gpurun_out/r02_bench_default.json 2> gpurun_out/r02_bench_default.err
echo "default bench rc=$? wall=$(( $(date +%s) - t0 )) s" | tee gpurun_out/r02_bench_default.wall
python - <<PY
import json
d=json.loads(open('gpurun_out/r02_bench_default.json').read().strip().splitlines()[-1])
print(round(d['ms_per_step'],3), round(d['value'],1), 'e2e', round(d['e2e']['value'],1), d['clocks'], 'cpu', d['cpu_baseline'])
print(json.dumps(d['config'].get('parity'))[:1500])
print(json.dumps(d.get('extra'))[:3500])
PY
