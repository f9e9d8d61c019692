#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
python -m pytest tests/test_gpu_entropy.py tests/test_gpu_checker.py tests/test_gpu_icip.py tests/test_gpu_model.py -m gpu -x -q 2>&1 | tail -n 15 | tee gpurun_out/pytest_gc.log
for c in 5 10 20; do echo "== v6 packed, $c CTAs/SM"; B200VC_GC_CTAS_PER_SM=$c python tools/gc_time.py 2>&1 | grep bits-only | tee gpurun_out/gc_time_v6_$c.log; done
echo "== v6 scalar math, 5 CTAs/SM"; B200VC_LIB=$PWD/tools/_bin/libb200vc_scalar.so python tools/gc_time.py 2>&1 | grep bits-only | tee gpurun_out/gc_time_v6_scalar.log
echo "== v6 scalar math, 10 CTAs/SM"; B200VC_GC_CTAS_PER_SM=10 B200VC_LIB=$PWD/tools/_bin/libb200vc_scalar.so python tools/gc_time.py 2>&1 | grep bits-only | tee gpurun_out/gc_time_v6_scalar10.log
