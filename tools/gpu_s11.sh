#!/bin/bash
# round 2: evidence run with the final defaults (pair kernel): tests, full bench line, ncu launch list / full set / traffic
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q --timeout 900 > gpurun_out/r2s11_tests.log 2>&1
echo "pytest rc=$?" >> gpurun_out/r2s11_tests.log
timeout 900 python bench.py --steps 10 --warmup 3 > gpurun_out/r2s11_bench.json 2> gpurun_out/r2s11_bench.err
echo "bench rc=$?" >> gpurun_out/r2s11_bench.err
Q="--no-e2e --no-cpu-baseline --no-beam --no-cfg4 --no-libc --no-variants --no-parity"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r2s11_launches.csv \
  python bench.py --steps 2 --warmup 1 $Q > gpurun_out/r2s11_ncu_bench.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:mf_tc_pair_kernel -s 1 -c 1 -o gpurun_out/r2s11_mf_prof \
  python bench.py --steps 1 --warmup 1 --templates 96 --samples 2000000 $Q > gpurun_out/r2s11_ncu_mf.log 2>&1
timeout 600 ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --clock-control none -k regex:mf_tc_pair_kernel -s 1 -c 1 \
  --csv --log-file gpurun_out/r2s11_traffic.csv python bench.py --steps 1 --warmup 1 $Q > gpurun_out/r2s11_ncu_traffic.log 2>&1
tail -c 400 gpurun_out/r2s11_tests.log
