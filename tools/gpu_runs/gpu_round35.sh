#!/bin/bash
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
timeout 600 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1
echo "smoke rc=$?"; tail -3 gpurun_out/smoke.log
timeout 600 python tools/decode_bench.py > gpurun_out/r01_decode_fp32_v12.json 2> gpurun_out/decode.err
echo "decode rc=$?"; cut -c1-600 gpurun_out/r01_decode_fp32_v12.json
timeout 600 python tools/decode_bench.py --precision bf16 > gpurun_out/r01_decode_bf16_v12.json 2>> gpurun_out/decode.err
echo "decode bf16 rc=$?"; cut -c1-300 gpurun_out/r01_decode_bf16_v12.json
