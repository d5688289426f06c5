#!/bin/bash
# round 2, GPU session 3: tests after compaction / mixed precision / host pool, bench, ncu evidence
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q --timeout 900 > gpurun_out/r2s3_tests.log 2>&1
echo "pytest rc=$?" >> gpurun_out/r2s3_tests.log
timeout 900 python bench.py --steps 10 --warmup 3 --no-cfg4 > gpurun_out/r2s3_bench.json 2> gpurun_out/r2s3_bench.err
echo "bench rc=$?" >> gpurun_out/r2s3_bench.err
Q="--no-e2e --no-cpu-baseline --no-beam --no-cfg4 --no-libc --no-variants --no-parity"
# launch list of the headline step (shares only)
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r2s3_launches.csv \
  python bench.py --steps 2 --warmup 1 $Q > gpurun_out/r2s3_ncu_bench.log 2>&1
# full set on the main kernel, reduced workload (as r1: 96 templates x 48 ch x 2e6 samples)
timeout 900 ncu --set full --clock-control none --import-source on -k regex:mf_tc_kernel -s 1 -c 1 -o gpurun_out/r2s3_mf_prof \
  python bench.py --steps 1 --warmup 1 --templates 96 --samples 2000000 $Q > gpurun_out/r2s3_ncu_mf.log 2>&1
# DRAM traffic of the full cfg2 launch
timeout 600 ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --clock-control none -k regex:mf_tc_kernel -s 1 -c 1 \
  --csv --log-file gpurun_out/r2s3_traffic.csv python bench.py --steps 1 --warmup 1 $Q > gpurun_out/r2s3_ncu_traffic.log 2>&1
tail -c 700 gpurun_out/r2s3_tests.log
