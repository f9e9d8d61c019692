#!/bin/bash
cd "$(dirname "$0")/.."
python -m pytest tests/test_gpu_flowguided.py tests/test_gpu_checker.py tests/test_gpu_icip.py -m gpu -q -s -p no:cacheprovider > gpurun_out/pytest_r2c.log 2>&1; tail -15 gpurun_out/pytest_r2c.log
python bench.py --steps 3 --warmup 3 > gpurun_out/bench_r2_1gpu_b.json 2> gpurun_out/bench_r2_1gpu_b.err; tail -c 600 gpurun_out/bench_r2_1gpu_b.json; tail -n 3 gpurun_out/bench_r2_1gpu_b.err
python bench.py --workload icip_gop16 --steps 1 --warmup 3 > gpurun_out/bench_r2_icip_1gpu.json 2> gpurun_out/bench_r2_icip_1gpu.err; tail -c 900 gpurun_out/bench_r2_icip_1gpu.json; tail -n 5 gpurun_out/bench_r2_icip_1gpu.err
