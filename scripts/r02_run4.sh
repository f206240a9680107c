#!/bin/bash
mkdir -p gpurun_out
IPCR_BENCH_TRACE=1 timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 5 --warmup 3 --no-e2e 2>gpurun_out/r02_b2.err >/dev/null
grep trace gpurun_out/r02_b2.err | tail -6
