#!/bin/bash
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
Q="--no-e2e --no-cpu-baseline --no-beam --no-cfg4 --no-libc --no-variants"
for pair in 1 0 1 0; do
  BPMF_TC_PAIR=$pair timeout 300 python bench.py --steps 6 --warmup 3 $Q >> gpurun_out/r2s10_pair$pair.json 2>> gpurun_out/r2s10_ab.err
done
BPMF_TC_PAIR=1 BPMF_TC_DEBUG=1 timeout 300 python bench.py --steps 6 --warmup 3 $Q --no-parity >> gpurun_out/r2s10_pair1_mmaonly.json 2>> gpurun_out/r2s10_ab.err
BPMF_TC_PAIR=0 BPMF_TC_DEBUG=1 timeout 300 python bench.py --steps 6 --warmup 3 $Q --no-parity >> gpurun_out/r2s10_pair0_mmaonly.json 2>> gpurun_out/r2s10_ab.err
timeout 900 python -m pytest tests/test_mf_gpu.py tests/test_full_size_gpu.py tests/test_call_patterns_gpu.py -m gpu -q --timeout 600 > gpurun_out/r2s10_tests.log 2>&1
tail -c 300 gpurun_out/r2s10_tests.log
