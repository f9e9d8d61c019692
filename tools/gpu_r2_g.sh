#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
run() { # tag, args...
tag=$1; shift
timeout 600 python bench.py "$@" --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/bench_r2_$tag.json 2> gpurun_out/bench_r2_$tag.err; tail -n 2 gpurun_out/bench_r2_$tag.err
python - <<PY
import json
try:
    d=json.load(open('gpurun_out/bench_r2_$tag.json'))
    print('$tag value',d['value'],'e2e',d['e2e']['value'],'ms',d['ms_per_step'],'roofline',d['roofline']['frac'], 'share', d['roofline'].get('hot_path_share_of_step'))
    for k,v in d['kernels'].items(): print(' ',k,v['launches'],round(v['ms_per_step'],3),round(v['frac_of_peak'],3))
except Exception as e: print('$tag failed', e)
PY
nvidia-smi --query-gpu=memory.used --format=csv,noheader
}
run G8 --gops-per-step 8
run G4_tf32 --gops-per-step 4 --conv-tf32
run G2 --gops-per-step 2
