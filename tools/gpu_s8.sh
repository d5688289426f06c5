#!/bin/bash
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q --timeout 900 > gpurun_out/r2s8_tests.log 2>&1
echo "pytest rc=$?" >> gpurun_out/r2s8_tests.log
Q="--no-e2e --no-cpu-baseline --no-beam --no-cfg4 --no-libc --no-variants --no-parity"
for dbg in 0 64 128 256 320 0 64; do
  BPMF_TC_DEBUG=$dbg timeout 300 python bench.py --steps 6 --warmup 3 $Q >> gpurun_out/r2s8_pol_$dbg.json 2>> gpurun_out/r2s8_pol.err
done
for dbg in 64 128 320; do
  BPMF_TC_DEBUG=$dbg timeout 600 ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --clock-control none -k regex:mf_tc_kernel -s 1 -c 1 \
    --csv --log-file gpurun_out/r2s8_traffic_$dbg.csv python bench.py --steps 1 --warmup 1 $Q > /dev/null 2>&1
done
for seg in 16 64; do
  BPMF_TC_SEG_TILES=$seg timeout 300 python bench.py --steps 6 --warmup 3 $Q >> gpurun_out/r2s8_seg_$seg.json 2>> gpurun_out/r2s8_pol.err
done
tail -c 400 gpurun_out/r2s8_tests.log
