#!/bin/bash
# flash attention bring-up: tests with and without the flash kernel, then microbench + bench
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_kernels_gpu.py -m gpu -q -k "attention" 2>&1 | tail -15 > gpurun_out/t_attn_flash.log
cat gpurun_out/t_attn_flash.log
WEEDCU_FLASH=0 timeout 600 python -m pytest tests/test_kernels_gpu.py -m gpu -q -k "attention" 2>&1 | tail -5 > gpurun_out/t_attn_noflash.log
cat gpurun_out/t_attn_noflash.log
timeout 1500 python -m pytest tests -m gpu -q 2>&1 | tail -8 > gpurun_out/t_gpu.log
cat gpurun_out/t_gpu.log
timeout 900 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/bench_full.log 2>&1
tail -n 1 gpurun_out/bench_full.log | cut -c1-300
WEEDCU_FLASH=0 timeout 900 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/bench_noflash.log 2>&1
tail -n 1 gpurun_out/bench_noflash.log | cut -c1-300
timeout 900 python tools/microbench.py > gpurun_out/microbench.log 2>&1
grep -E "softmax|layernorm|gemm" gpurun_out/microbench.log | head -50
