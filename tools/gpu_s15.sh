#!/bin/bash
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_multigpu.py -m gpu -q --timeout 500 > gpurun_out/r2s15_tests.log 2>&1
echo "pytest rc=$?" >> gpurun_out/r2s15_tests.log
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29521 \
  bench.py --gpus 2 --steps 5 --warmup 3 > gpurun_out/r2s15_bench_n2.json 2> gpurun_out/r2s15_bench_n2.err
echo "bench rc=$?" >> gpurun_out/r2s15_bench_n2.err
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29522 \
  bench.py --impl reference --gpus 2 --steps 1 --warmup 1 > gpurun_out/r2s15_ref_n2.json 2> gpurun_out/r2s15_ref_n2.err
tail -c 300 gpurun_out/r2s15_tests.log; tail -c 300 gpurun_out/r2s15_bench_n2.err
