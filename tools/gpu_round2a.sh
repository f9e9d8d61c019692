#!/bin/bash
# round-2 GPU pass A: acceptance + flex + warp tests, warp2 timing, first bench lines
cd "$(dirname "$0")/.."
python -m pytest tests/test_gpu_acceptance.py tests/test_gpu_flexrate.py tests/test_gpu_warp.py tests/test_gpu_coding.py tests/test_gpu_model.py -m gpu -q -s -p no:cacheprovider > gpurun_out/pytest_r2b.log 2>&1
tail -25 gpurun_out/pytest_r2b.log
python tools/warp2_time.py > gpurun_out/warp2_time_v2.log 2>&1; B200VC_WARP2_V2=0 python tools/warp2_time.py > gpurun_out/warp2_time_v1.log 2>&1
tail -4 gpurun_out/warp2_time_v2.log gpurun_out/warp2_time_v1.log
python bench.py --steps 3 --warmup 3 > gpurun_out/bench_r2_1gpu_a.json 2> gpurun_out/bench_r2_1gpu_a.err; tail -c 1500 gpurun_out/bench_r2_1gpu_a.json; tail -3 gpurun_out/bench_r2_1gpu_a.err
