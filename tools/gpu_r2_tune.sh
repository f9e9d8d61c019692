#!/bin/bash
cd "$(dirname "$0")/.."
timeout 200 python tools/ab_parity.py 2>&1 | tail -n 1
B200VC_LIB=$PWD/tools/_bin/libb200vc_prev.so timeout 200 python tools/ab_parity.py 2>&1 | tail -n 1
run() { tag=$1; shift
( time timeout 700 python bench.py "$@" --steps 2 --warmup 3 --no-e2e --no-cpu-baseline --no-tf32-leg > gpurun_out/tune_$tag.json 2> gpurun_out/tune_$tag.err ) 2>&1 | grep real; grep "\[bench\]" gpurun_out/tune_$tag.err | tr '\n' ';'; python -c "
import json; d=json.load(open('gpurun_out/tune_$tag.json')); print('$tag', round(d['value'],3), 'B-frames/s', round(d['ms_per_step'],1),'ms')"
}
run lim6_G8 --cudnn-benchmark-limit 6
