#!/bin/bash
# 8 GPUs: coalesced small-gradient exchange + split Adam at N = 8; --check-dp at N = 2 on the same box
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
run8() { # name, env...
  name=$1; shift
  env "$@" timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29800 bench.py --gpus 8 --steps 10 --warmup 3 --no-cpu-baseline --no-parity --no-configs --no-clocks > gpurun_out/r02_dp8_$name.json 2> gpurun_out/r02_dp8_$name.err
  python - <<PY
import json
try:
    d=json.loads(open('gpurun_out/r02_dp8_$name.json').read().strip().splitlines()[-1])
    print('$name', round(d['ms_per_step'],3), round(d['value'],1))
except Exception as e: print('$name','ERR',e)
PY
}
run8 coalesce_split A=1
run8 coalesce_nosplit WH_DP_SPLIT_ADAM=0
CUDA_VISIBLE_DEVICES=0,1 timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29611 bench.py --gpus 2 --check-dp > gpurun_out/r02_check_dp_n2.json 2> gpurun_out/r02_check_dp_n2.err
echo "check-dp rc=$?"
python - <<PY
import json
d=json.loads(open('gpurun_out/r02_check_dp_n2.json').read().strip().splitlines()[-1])
print(d['ok'])
for k in ('fp32','bf16','bf16_reference_layernorm_chain'):
    r=d[k]; print(' ',k, 'loss', r['loss_rel_diff'], 'm', r['adam_m_rel_to_max_diff'], 'v', r['adam_v_rel_to_max_diff'], r['ok'])
PY
