#!/bin/bash
# 2 GPUs: does leaving SMs to NCCL help the overlapped exchange?
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
run2() { name=$1; shift
  env "$@" timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29800 bench.py --gpus 2 --steps 10 --warmup 3 --no-cpu-baseline --no-parity --no-configs --no-clocks > gpurun_out/r02_dp2_$name.json 2> gpurun_out/r02_dp2_$name.err
  python - <<PY
import json
try:
    d=json.loads(open('gpurun_out/r02_dp2_$name.json').read().strip().splitlines()[-1])
    print('$name', round(d['ms_per_step'],3), round(d['value'],1))
except Exception as e: print('$name','ERR',e)
PY
}
run2 default A=1
run2 sms140_nch8 WEEDCU_GEMM_SMS=140 NCCL_MAX_NCHANNELS=8
run2 sms132_nch16 WEEDCU_GEMM_SMS=132 NCCL_MAX_NCHANNELS=16
run2 sms144_nch4 WEEDCU_GEMM_SMS=144 NCCL_MAX_NCHANNELS=4
run2 nooverlap WH_DP_OVERLAP=0
WEEDCU_GEMM_SMS=140 timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-parity --no-configs --no-clocks > gpurun_out/r02_n1_sms140.json 2>/dev/null
python - <<PY
import json
d=json.loads(open('gpurun_out/r02_n1_sms140.json').read().strip().splitlines()[-1]); print('n1 sms140', round(d['ms_per_step'],3))
PY
