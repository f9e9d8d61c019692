#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
# tiny dry run of the whole bench flow first (a crash in a late leg shows up in seconds, not minutes)
timeout 300 python bench.py --height 192 --width 256 --steps 1 --warmup 3 --gops-per-step 2 > gpurun_out/bench_tiny.json 2> gpurun_out/bench_tiny.err || { echo TINY FAILED; tail -n 12 gpurun_out/bench_tiny.err; exit 1; }
head -c 300 gpurun_out/bench_tiny.json; echo
( time timeout 900 python bench.py --steps 3 --warmup 3 > gpurun_out/bench_r2g_lhbdc.json 2> gpurun_out/bench_r2g_lhbdc.err ) 2>&1 | grep real; grep "\[bench\]" gpurun_out/bench_r2g_lhbdc.err
python - <<PY
import json
d=json.load(open('gpurun_out/bench_r2g_lhbdc.json'))
print(d['value'], d['e2e'], d['conv_tf32'], d['cudnn'], d['roofline']['frac'], d['cpu_baseline'])
print(d['parity'])
for k,v in d['kernels'].items(): print(' ',k,v['launches'],round(v['ms_per_step'],3),round(v['frac_of_peak'],3))
PY
timeout 240 compute-sanitizer --tool memcheck python tools/sanitize_gc.py 2>&1 | tail -n 4 | tee gpurun_out/sanitize_gc.log
