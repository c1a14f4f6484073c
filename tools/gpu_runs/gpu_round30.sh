#!/bin/bash
# 2-GPU data-parallel bench (driver-style launch) next to the same-box N = 1 run
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
timeout 600 python bench.py --gpus 1 --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/r01_bench_n1_v14.json 2> gpurun_out/n1.err
echo "n1 rc=$?"; cut -c1-260 gpurun_out/r01_bench_n1_v14.json
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 10 --warmup 3 > gpurun_out/r01_bench_n2_v14.json 2> gpurun_out/n2.err
echo "n2 rc=$?"; cut -c1-260 gpurun_out/r01_bench_n2_v14.json; tail -3 gpurun_out/n2.err
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --impl reference --gpus 2 --steps 1 --warmup 0 > gpurun_out/r01_bench_ref_n2_v14.json 2> gpurun_out/refn2.err
echo "ref n2 rc=$?"; cut -c1-200 gpurun_out/r01_bench_ref_n2_v14.json
