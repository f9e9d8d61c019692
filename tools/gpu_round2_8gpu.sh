#!/bin/bash
# 8-GPU lines: strong scaling of ONE 601-frame sequence (75 GOP-8 sharded 10/9), Flex all-quality (8 units), OJSP segments
cd "$(dirname "$0")/.."
run() { python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29611 bench.py --gpus 8 "$@"; }
run --sequence-frames 601 --steps 1 --warmup 3 > gpurun_out/bench_r2_strong601_8gpu.json 2> gpurun_out/bench_r2_strong601_8gpu.err; tail -c 400 gpurun_out/bench_r2_strong601_8gpu.json; tail -n 2 gpurun_out/bench_r2_strong601_8gpu.err
run --workload flex_gop16_allq --steps 1 --warmup 3 > gpurun_out/bench_r2_flex_8gpu.json 2> gpurun_out/bench_r2_flex_8gpu.err; tail -c 300 gpurun_out/bench_r2_flex_8gpu.json; tail -n 2 gpurun_out/bench_r2_flex_8gpu.err
run --workload ojsp_search_4k --steps 3 --warmup 3 > gpurun_out/bench_r2_ojsp_8gpu.json 2> gpurun_out/bench_r2_ojsp_8gpu.err; tail -c 300 gpurun_out/bench_r2_ojsp_8gpu.json; tail -n 2 gpurun_out/bench_r2_ojsp_8gpu.err
