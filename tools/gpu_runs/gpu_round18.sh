#!/bin/bash
# ncu --set full captures of the step's main kernels (kept under the 64 MiB return limit: no source import, 22 launches)
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
timeout 900 ncu --set full --clock-control none -k regex:"gemm_bf16_kernel|flash_attn|layernorm_fwd_stats|layernorm_fwd_apply|layernorm_bwd_rows|layernorm_bwd_apply|ce_bwd_pack|ce_fwd_partial|adam_multi|gelu_fwd_bf16|gelu_grad_pack|heads_pack_vec|pack_bf16_colsum" --launch-skip 1020 -c 22 -f -o gpurun_out/r01_full_final python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-clocks > gpurun_out/ncu_full.log 2>&1
tail -2 gpurun_out/ncu_full.log | cut -c1-200
ls -la gpurun_out/*.ncu-rep
