#!/bin/bash
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29531 \
  bench.py --gpus 8 --steps 6 --warmup 3 > gpurun_out/r2s17_bench_n8.json 2> gpurun_out/r2s17_bench_n8.err
echo "bench rc=$?" >> gpurun_out/r2s17_bench_n8.err
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29532 \
  bench.py --gpus 4 --steps 6 --warmup 3 --no-cfg4 --no-beam > gpurun_out/r2s17_bench_n4.json 2> gpurun_out/r2s17_bench_n4.err
echo "bench rc=$?" >> gpurun_out/r2s17_bench_n4.err
tail -c 300 gpurun_out/r2s17_bench_n8.err; tail -c 300 gpurun_out/r2s17_bench_n4.err
