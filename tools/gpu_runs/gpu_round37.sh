#!/bin/bash
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
timeout 900 python -m pytest tests -x -q -m gpu -k "argmax or layernorm or skinny or decode or greedy or kv_cache or encoder or residual" > gpurun_out/decode_tests.log 2>&1
echo "tests rc=$?"; tail -3 gpurun_out/decode_tests.log
timeout 600 python tools/decode_bench.py > gpurun_out/r01_decode_fp32_v13.json 2> gpurun_out/decode.err
echo "decode rc=$?"; python - <<'PY'
import json
d=json.loads(open('gpurun_out/r01_decode_fp32_v13.json').read().strip().splitlines()[-1])
print(d['value'], d['ms_per_step'], d['launches_per_step'], d['roofline']['frac'])
for k,v in d['kernel_breakdown'].items(): print(k, round(v['ms_per_step'],4), v['launches_per_step'])
PY
