#!/bin/bash
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
./tools/red_bench > gpurun_out/r2s14_red_bench.txt 2>&1
for pat in red.v2 red.v4 st.v4 red.u64 red.f32; do
  nvidia-smi --query-gpu=clocks.sm,power.draw --format=csv,noheader,nounits -lms 100 > gpurun_out/r2s14_power_$pat.csv &
  SMI=$!
  sleep 0.5
  ./tools/red_bench $pat 5 | grep SUSTAINED >> gpurun_out/r2s14_red_bench.txt
  kill $SMI
done
# idle power for reference
nvidia-smi --query-gpu=clocks.sm,power.draw --format=csv,noheader,nounits -lms 100 > gpurun_out/r2s14_power_idle.csv &
SMI=$!; sleep 2; kill $SMI
cat gpurun_out/r2s14_red_bench.txt
