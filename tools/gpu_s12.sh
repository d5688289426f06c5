#!/bin/bash
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
Q="--no-e2e --no-cpu-baseline --no-beam --no-cfg4 --no-libc --no-variants --no-parity"
for rep in 1 2; do
for pair in 1 0; do
for seg in 16 32 64; do
  BPMF_TC_PAIR=$pair BPMF_TC_SEG_TILES=$seg timeout 300 python bench.py --steps 6 --warmup 3 $Q >> gpurun_out/r2s12_p${pair}_s${seg}.json 2>> gpurun_out/r2s12.err
done; done; done
for pair in 1 0; do
for seg in 16 32; do
  K=mf_tc_kernel; [ $pair = 1 ] && K=mf_tc_pair_kernel
  BPMF_TC_PAIR=$pair BPMF_TC_SEG_TILES=$seg timeout 600 ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --clock-control none -k regex:$K -s 1 -c 1 \
    --csv --log-file gpurun_out/r2s12_traffic_p${pair}_s${seg}.csv python bench.py --steps 1 --warmup 1 $Q > /dev/null 2>&1
done; done
echo done
