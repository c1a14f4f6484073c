#!/bin/bash
# full GPU tests (no -x) + ncu --set full captures of the attention / LayerNorm / CE kernels inside the bench step
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q 2>&1 | tail -15 > gpurun_out/t_gpu.log
cat gpurun_out/t_gpu.log
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"flash_attn|heads_pack|layernorm_fwd_reg|unheads" --launch-skip 20 -c 5 -f -o gpurun_out/r01_full_attn_ln python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-clocks > gpurun_out/ncu1.log 2>&1
tail -3 gpurun_out/ncu1.log | cut -c1-200
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"layernorm_bwd_reg|ce_fwd_partial|ce_bwd|pack_bf16_colsum|gemm_bf16_kernel<192, 4, 1, 0>" --launch-skip 4 -c 8 -f -o gpurun_out/r01_full_bwd python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-clocks > gpurun_out/ncu2.log 2>&1
tail -3 gpurun_out/ncu2.log | cut -c1-200
ls -la gpurun_out/*.ncu-rep
