#!/bin/bash
# 8 GPUs: data-parallel knobs at N = 8 (10 steps each)
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
run8 default A=1
run8 chain WH_DP_CHAIN_ADAM=1
run8 bb16m WH_DP_BUCKET_BYTES=16777216
run8 bb64m WH_DP_BUCKET_BYTES=67108864
run8 nch8 NCCL_MAX_NCHANNELS=8
run8 nch16 NCCL_MAX_NCHANNELS=16
run8 nooverlap WH_DP_OVERLAP=0
