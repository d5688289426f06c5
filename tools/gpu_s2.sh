#!/bin/bash
# round 2, GPU session 2 (2 GPUs): multi-GPU tests + bench at N=2
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
nvidia-smi topo -m > gpurun_out/r2s2_topo.txt 2>&1
timeout 600 python -m pytest tests/test_multigpu.py tests/test_libc_gpu.py -m gpu -q --timeout 500 > gpurun_out/r2s2_tests.log 2>&1
echo "pytest rc=$?" >> gpurun_out/r2s2_tests.log
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 \
  bench.py --gpus 2 --steps 5 --warmup 3 > gpurun_out/r2s2_bench_n2.json 2> gpurun_out/r2s2_bench_n2.err
echo "bench rc=$?" >> gpurun_out/r2s2_bench_n2.err
tail -c 600 gpurun_out/r2s2_tests.log; tail -c 800 gpurun_out/r2s2_bench_n2.err
