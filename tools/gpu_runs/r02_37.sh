#!/bin/bash
# 8 GPUs: --check-dp (global batch 16 = 2 per rank... use --batch 2 -> GB 16)
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
timeout 700 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29611 bench.py --gpus 8 --check-dp --batch 2 > gpurun_out/r02_check_dp_n8.json 2> gpurun_out/r02_check_dp_n8.err
echo "check-dp rc=$?"
python - <<PY
import json
d=json.loads(open('gpurun_out/r02_check_dp_n8.json').read().strip().splitlines()[-1])
print(d['ok'], d['n_gpus'], d['global_batch'], d['batch_per_gpu'])
for k in ('fp32','bf16','bf16_reference_layernorm_chain'):
    r=d[k]; print(' ',k, 'loss', r['loss_rel_diff'], 'm', r['adam_m_rel_to_max_diff'], 'v', r['adam_v_rel_to_max_diff'], r['ok'])
PY
tail -3 gpurun_out/r02_check_dp_n8.err
