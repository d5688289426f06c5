#!/bin/bash
# round 2, GPU session 1: tests + epilogue A/B + full bench line
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.limit --format=csv > gpurun_out/r2s1_gpu.txt 2>&1
nproc >> gpurun_out/r2s1_gpu.txt; grep -m1 "model name" /proc/cpuinfo >> gpurun_out/r2s1_gpu.txt; free -g | head -2 >> gpurun_out/r2s1_gpu.txt
Q="--no-e2e --no-cpu-baseline --no-beam --no-cfg4 --no-libc --no-variants --no-parity"
for epi in 0 1 0 1; do
  BPMF_TC_EPI=$epi timeout 300 python bench.py --steps 6 --warmup 3 $Q >> gpurun_out/r2s1_ab_epi$epi.json 2>> gpurun_out/r2s1_ab.err
done
timeout 1500 python -m pytest tests -m gpu -q --timeout 900 > gpurun_out/r2s1_tests.log 2>&1
echo "pytest rc=$?" >> gpurun_out/r2s1_tests.log
timeout 900 python bench.py --steps 10 --warmup 3 > gpurun_out/r2s1_bench.json 2> gpurun_out/r2s1_bench.err
echo "bench rc=$?" >> gpurun_out/r2s1_bench.err
tail -c 1500 gpurun_out/r2s1_tests.log
