#!/bin/bash
# last confirmation of the committed tree after a clean rebuild: GPU suite, smoke(), one short bench
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -x -q 2>&1 | tail -2 | tee gpurun_out/r02_gpu_tests_48.log
timeout 200 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
timeout 200 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-parity --no-configs > gpurun_out/r02_bench_48.json 2>/dev/null
python - <<PY
import json
d=json.loads(open('gpurun_out/r02_bench_48.json').read().strip().splitlines()[-1])
print(round(d['ms_per_step'],3), round(d['value'],1), 'e2e', round(d['e2e']['value'],1), d['clocks']['reasons'])
PY
