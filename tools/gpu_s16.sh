#!/bin/bash
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
# memcheck over the new kernels: shuffle-free epilogue + per-channel tiles + mixed precision, pair kernel, peaks, gated detection
timeout 900 compute-sanitizer --tool memcheck --error-exitcode 1 python -m pytest tests/test_mf_gpu.py -m gpu -q -k "cfg1 or edge_shapes or auto_policy or bounds" > gpurun_out/r2s16_memcheck_mf.log 2>&1
echo "rc=$?" >> gpurun_out/r2s16_memcheck_mf.log
BPMF_TC_PAIR=1 timeout 900 compute-sanitizer --tool memcheck --error-exitcode 1 python -m pytest tests/test_mf_gpu.py -m gpu -q -k "cfg1 or edge_shapes" > gpurun_out/r2s16_memcheck_pair.log 2>&1
echo "rc=$?" >> gpurun_out/r2s16_memcheck_pair.log
timeout 900 compute-sanitizer --tool memcheck --error-exitcode 1 python -m pytest tests/test_peaks.py tests/test_device_gpu.py tests/test_libc_gpu.py -m gpu -q -k "not day_length" > gpurun_out/r2s16_memcheck_misc.log 2>&1
echo "rc=$?" >> gpurun_out/r2s16_memcheck_misc.log
for f in mf pair misc; do tail -4 gpurun_out/r2s16_memcheck_$f.log; done
