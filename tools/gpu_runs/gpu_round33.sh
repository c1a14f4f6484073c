#!/bin/bash
# leave-one-class-out under programmatic dependent launch: which class must be plain for the loss to return to 10.8826?
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
for cls in none 0 1 3 4 6 7 8 10 11 13; do
  if [ "$cls" = "none" ]; then mask=-1; else mask=$(( -1 ^ (1 << cls) )); fi
  WEEDCU_PDL=1 WEEDCU_PDL_CLASSES=$mask timeout 90 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/pdl_loo$cls.json 2> gpurun_out/pdl_loo$cls.err
  python - <<PY
import json
try:
    d=json.loads(open('gpurun_out/pdl_loo$cls.json').read().strip().splitlines()[-1]); print('without class $cls', round(d['ms_per_step'],3), d['config']['loss_last'])
except Exception as e: print('without class $cls no result', e)
PY
done
