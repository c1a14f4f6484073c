#!/bin/bash
# ncu --set full of the step's main kernels, round-2 code: (A) two forward layers, (B) LM head / loss / first backward layer
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
RX='regex:gemm_bf16_pair_kernel|gemm_bf16_kernel|flash_attn|gelu_grad_pack|ce_bwd_pack|ce_fwd_partial|ce_fwd_stats|adam_multi|layernorm_fwd_stats|layernorm_fwd_apply|layernorm_stats_merge|layernorm_bwd_apply|layernorm_bwd_rows|heads_pack|pack_bf16_colsum'
timeout 600 ncu --set full --clock-control none -k "$RX" --launch-skip 773 -c 24 -f -o gpurun_out/r02_full_fwd python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-clocks --no-parity --no-configs > gpurun_out/r02_full_fwd.log 2>&1
echo "A rc=$?"
timeout 600 ncu --set full --clock-control none -k "$RX" --launch-skip 874 -c 40 -f -o gpurun_out/r02_full_bwd python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-clocks --no-parity --no-configs > gpurun_out/r02_full_bwd.log 2>&1
echo "B rc=$?"; ls -la gpurun_out/r02_full_*.ncu-rep
