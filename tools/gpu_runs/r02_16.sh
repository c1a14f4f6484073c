#!/bin/bash
# 2 GPUs: --check-dp at global batch 8 (4 per rank) and global batch 16, 1 step and 3 steps
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
for b in 4 8; do
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 2961$b bench.py --gpus 2 --check-dp --batch $b > gpurun_out/r02_check_dp_n2_b$b.json 2> gpurun_out/r02_check_dp_n2_b$b.err
echo "check-dp b=$b rc=$?"
python - <<PY
import json
d=json.loads(open('gpurun_out/r02_check_dp_n2_b$b.json').read().strip().splitlines()[-1])
for k in ('fp32','bf16'):
    r=d[k]; print(k, 'loss', r['loss_rel_diff'], 'm', r['adam_m_rel_to_max_diff'], 'v', r['adam_v_rel_to_max_diff'], 'gated', r['params_gated']); print(r['worst_params'])
PY
done
