#!/bin/bash
# end-of-session evidence: full GPU suite, bench (N=1, cpu_baseline), reference arm, ncu launch list + GEMM DRAM traffic + --set full
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -x -q -m gpu > gpurun_out/gpu_tests_final.log 2>&1
echo "gpu tests rc=$?"; tail -3 gpurun_out/gpu_tests_final.log
timeout 900 python bench.py --steps 20 --warmup 5 > gpurun_out/r01_bench_full_v16.json 2> gpurun_out/bench_v16.err
echo "bench rc=$?"; cut -c1-300 gpurun_out/r01_bench_full_v16.json
timeout 600 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/r01_bench_reference_v16.json 2> gpurun_out/bench_ref_v16.err
echo "ref rc=$?"
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --launch-skip 1050 -c 700 --csv --log-file gpurun_out/r01_launches_v16.csv python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-clocks > gpurun_out/bench_ncu.log 2>&1
echo "launch list rc=$?"
timeout 900 ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --clock-control none -k regex:gemm_bf16 --launch-skip 333 -c 111 --csv --log-file gpurun_out/r01_gemm_dram_v16.csv python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-clocks > gpurun_out/bench_ncu2.log 2>&1
echo "gemm dram rc=$?"
timeout 900 ncu --set full --clock-control none -k regex:"gemm_bf16_pair_kernel|gemm_bf16_kernel|flash_attn|gelu_grad_pack|gelu_fwd_bf16|ce_bwd_pack|ce_fwd_partial|adam_multi|layernorm_fwd_stats|layernorm_fwd_apply|layernorm_bwd_apply|layernorm_bwd_rows|heads_pack|pack_bf16_colsum" --launch-skip 1100 -c 30 -f -o gpurun_out/r01_full_v16 python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-clocks > gpurun_out/ncu_full.log 2>&1
echo "ncu full rc=$?"; ls -la gpurun_out/r01_full_v16.ncu-rep
timeout 900 python tools/microbench.py --group ew --out gpurun_out/r01_microbench_ew_v16.json > gpurun_out/r01_microbench_ew_v16.log 2>&1
echo "microbench rc=$?"; tail -5 gpurun_out/r01_microbench_ew_v16.log
timeout 600 python tools/microbench.py --group tune --out gpurun_out/r01_tune_v16.json > gpurun_out/r01_tune_v16.log 2>&1
echo "tune rc=$?"; grep -E "mode        0 |BEST" gpurun_out/r01_tune_v16.log | head -32
timeout 600 python tools/decode_bench.py > gpurun_out/r01_decode_fp32_v16.json 2> gpurun_out/decode.err
echo "decode rc=$?"; cut -c1-200 gpurun_out/r01_decode_fp32_v16.json
