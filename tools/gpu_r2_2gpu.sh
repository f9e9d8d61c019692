#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
( time timeout 800 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus 2 --steps 2 --warmup 3 --no-tf32-leg > gpurun_out/bench_r2g_2gpu.json 2> gpurun_out/bench_r2g_2gpu.err ) 2>&1 | grep real; grep "\[bench\]" gpurun_out/bench_r2g_2gpu.err; head -c 400 gpurun_out/bench_r2g_2gpu.json; echo; tail -n 3 gpurun_out/bench_r2g_2gpu.err
