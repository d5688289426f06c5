#!/bin/bash
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
Q="--no-e2e --no-cpu-baseline --no-beam --no-cfg4 --no-libc --no-variants --no-parity"
for ns in 7 3 5 7; do
BPMF_GATHER_STREAMS=$ns timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 2954$ns \
  bench.py --gpus 8 --steps 6 --warmup 3 $Q >> gpurun_out/r2s18_n8_streams$ns.json 2>> gpurun_out/r2s18.err
done
BPMF_GATHER_STREAMS=7 timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29549 \
  bench.py --gpus 8 --steps 6 --warmup 3 $Q --batch 32 >> gpurun_out/r2s18_n8_streams7_b32.json 2>> gpurun_out/r2s18.err
tail -c 400 gpurun_out/r2s18.err
