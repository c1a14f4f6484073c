#!/bin/bash
# final evidence of the session: bench (N=1, with cpu_baseline), reference arm, ncu --set full of the step's main kernels, microbench
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
timeout 900 python bench.py --steps 20 --warmup 5 > gpurun_out/r01_bench_full_v10.json 2> gpurun_out/bench_v10.err
echo "bench rc=$?"; cut -c1-400 gpurun_out/r01_bench_full_v10.json
timeout 600 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/r01_bench_reference_v10.json 2> gpurun_out/bench_ref_v10.err
echo "ref rc=$?"; cut -c1-300 gpurun_out/r01_bench_reference_v10.json
timeout 900 ncu --set full --clock-control none -k regex:"gemm_bf16_pair_kernel|gemm_bf16_kernel|flash_attn|gelu_grad_pack|ce_bwd_pack|adam_multi|layernorm_fwd_stats|layernorm_bwd_apply" --launch-skip 1300 -c 24 -f -o gpurun_out/r01_full_v10 python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-clocks > gpurun_out/ncu_full.log 2>&1
echo "ncu rc=$?"; ls -la gpurun_out/r01_full_v10.ncu-rep
timeout 900 python tools/microbench.py --group gemm --out gpurun_out/r01_microbench_gemm_v10.json > gpurun_out/r01_microbench_gemm_v10.log 2>&1
tail -30 gpurun_out/r01_microbench_gemm_v10.log
