#!/bin/bash
# round 2, 8-GPU session: bench at N=8 (headline weak scaling + cfg4 strong scaling + beam sharded + e2e), cfg5 day
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
nvidia-smi topo -m > gpurun_out/r2s7_topo.txt 2>&1
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29512 \
  bench.py --gpus 8 --steps 6 --warmup 3 > gpurun_out/r2s7_bench_n8.json 2> gpurun_out/r2s7_bench_n8.err
echo "bench rc=$?" >> gpurun_out/r2s7_bench_n8.err
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29513 \
  tools/full_day_sharded.py > gpurun_out/r2s7_fullday_n8.json 2> gpurun_out/r2s7_fullday_n8.err
echo "fullday rc=$?" >> gpurun_out/r2s7_fullday_n8.err
tail -c 600 gpurun_out/r2s7_bench_n8.err; tail -c 600 gpurun_out/r2s7_fullday_n8.err; cat gpurun_out/r2s7_fullday_n8.json
