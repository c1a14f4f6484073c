#!/bin/bash
# programmatic dependent launch with wait-before-trigger: everything once with WEEDCU_PDL=1
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
export WEEDCU_PDL=1
timeout 900 python -m pytest tests -x -q -m gpu > gpurun_out/gpu_tests_pdl.log 2>&1
echo "gpu tests rc=$?"; tail -3 gpurun_out/gpu_tests_pdl.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke_pdl.log 2>&1
echo "smoke rc=$?"; tail -1 gpurun_out/smoke_pdl.log
for i in 1 2 3 4; do
  timeout 120 python bench.py --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/pdl_bench$i.json 2> gpurun_out/pdl_bench$i.err
  echo "bench $i rc=$?"; python - <<PY
import json
try:
    d=json.loads(open('gpurun_out/pdl_bench$i.json').read().strip().splitlines()[-1]); print(round(d['ms_per_step'],3), round(d['value'],1), round(d['e2e']['value'],1), d['config']['loss_last'])
except Exception as e: print('no result', e)
PY
done
timeout 300 python tools/decode_bench.py > gpurun_out/decode_pdl.json 2> gpurun_out/decode_pdl.err
echo "decode rc=$?"; cut -c1-220 gpurun_out/decode_pdl.json
timeout 600 python tools/microbench.py --group tune --out gpurun_out/tune_pdl.json > gpurun_out/tune_pdl.log 2>&1
echo "tune rc=$?"; grep -E "mode        0 " gpurun_out/tune_pdl.log | head -20
