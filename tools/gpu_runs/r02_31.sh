#!/bin/bash
# parity legs (bf16 vs fp32 loss trajectory, 5 steps) with the round-2 approximations toggled
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
for cfg in "bf16_act_grad=1" "bf16_act_grad=0" "bf16_act_grad=0,epilogue_stats=0"; do
WH_CONFIG="$cfg" timeout 600 python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-configs --no-clocks > gpurun_out/r02_parity_tmp.json 2> gpurun_out/r02_parity_tmp.err
python - <<PY
import json
d=json.loads(open('gpurun_out/r02_parity_tmp.json').read().strip().splitlines()[-1])
p=d['config']['parity']
print("$cfg", round(d['ms_per_step'],3), p['loss_rel_diff_bf16_vs_fp32'], [round(x,5) for x in p['loss_bf16']], [round(x,5) for x in p['loss_fp32']])
PY
done
