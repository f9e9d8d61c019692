#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_warp.py -m gpu -x -q 2>&1 | tail -n 8 | tee gpurun_out/pytest_w.log
echo "== packed"; timeout 120 python tools/warp2_time.py 2>&1 | tee gpurun_out/warp2_time_packed.log
echo "== previous build"; B200VC_LIB=$PWD/tools/_bin/libb200vc_prev.so timeout 120 python tools/warp2_time.py 2>&1 | tee gpurun_out/warp2_time_prev.log
