#!/bin/bash
# Round-1 evidence run: default bench line (with cpu_baseline), reference arm, ncu launch list of
# the bench command, and ncu --set full captures of the heaviest kernels.
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -x -q 2>&1 | tail -5 > gpurun_out/t_gpu.log
cat gpurun_out/t_gpu.log
S=$(date +%s)
timeout 900 python bench.py > gpurun_out/bench_default.log 2>&1
echo "default bench: $(( $(date +%s) - S )) s" > gpurun_out/timing.log
S=$(date +%s)
timeout 900 python bench.py --impl reference --steps 1 --warmup 0 > gpurun_out/bench_reference.log 2>&1
echo "reference arm: $(( $(date +%s) - S )) s" >> gpurun_out/timing.log
S=$(date +%s)
timeout 1200 ncu --metrics gpu__time_duration.sum --clock-control none --csv \
  --log-file gpurun_out/r01_launches.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-clocks \
  > gpurun_out/bench_under_ncu.log 2>&1
echo "ncu launch list: $(( $(date +%s) - S )) s" >> gpurun_out/timing.log
S=$(date +%s)
for k in gemm_bf16_kernel pack_bf16_kernel softmax_strided_fwd layernorm_bwd_kernel; do
  timeout 600 ncu --set full --clock-control none --import-source on -k regex:$k -s 4 -c 2 \
    -o gpurun_out/r01_full_$k -f python bench.py --steps 1 --warmup 0 --layers 2 --no-cpu-baseline --no-clocks \
    > gpurun_out/ncu_full_$k.log 2>&1
done
echo "ncu full: $(( $(date +%s) - S )) s" >> gpurun_out/timing.log
cat gpurun_out/timing.log
tail -n 1 gpurun_out/bench_default.log | cut -c1-1500
tail -n 1 gpurun_out/bench_reference.log | cut -c1-600
ls -la gpurun_out
