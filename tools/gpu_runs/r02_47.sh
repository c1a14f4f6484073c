#!/bin/bash
# ncu --set full of one forward layer and one backward layer of the step (round-2 code); only the text summaries travel back
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out /tmp/ncu
RX='regex:gemm_bf16_pair_kernel|gemm_bf16_kernel|flash_attn|gelu_grad_pack|ce_bwd_pack|ce_fwd_partial|ce_fwd_stats|adam_multi|layernorm_fwd_stats|layernorm_fwd_apply|layernorm_stats_merge|layernorm_bwd_apply|layernorm_bwd_rows|heads_pack|pack_bf16_colsum'
timeout 300 ncu --set full --clock-control none -k "$RX" --launch-skip 777 -c 12 -f -o /tmp/ncu/r02_full_fwd python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-clocks --no-parity --no-configs > /tmp/ncu/fwd.log 2>&1
echo "A rc=$?"
timeout 300 ncu --set full --clock-control none -k "$RX" --launch-skip 883 -c 14 -f -o /tmp/ncu/r02_full_bwd python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-clocks --no-parity --no-configs > /tmp/ncu/bwd.log 2>&1
echo "B rc=$?"
python tools/ncu_summary.py full /tmp/ncu/r02_full_fwd.ncu-rep > gpurun_out/r02_full_step_kernels_fwd.txt 2>&1
python tools/ncu_summary.py full /tmp/ncu/r02_full_bwd.ncu-rep > gpurun_out/r02_full_step_kernels_bwd.txt 2>&1
wc -l gpurun_out/r02_full_step_kernels_*.txt
