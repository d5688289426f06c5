#!/bin/bash
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q --timeout 900 > gpurun_out/r2s13_tests.log 2>&1
echo "pytest rc=$?" >> gpurun_out/r2s13_tests.log
timeout 900 python bench.py --steps 10 --warmup 3 > gpurun_out/r2s13_bench.json 2> gpurun_out/r2s13_bench.err
echo "bench rc=$?" >> gpurun_out/r2s13_bench.err
timeout 300 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/r2s13_ref.json 2> gpurun_out/r2s13_ref.err
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r2s13_smoke.txt 2>&1
tail -c 300 gpurun_out/r2s13_tests.log; cat gpurun_out/r2s13_smoke.txt | tail -2
