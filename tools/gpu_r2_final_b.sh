#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
( time timeout 900 python bench.py --steps 3 --warmup 3 > gpurun_out/bench_r2f_lhbdc.json 2> gpurun_out/bench_r2f_lhbdc.err ) 2>&1 | tail -n 4; grep "\[bench\]" gpurun_out/bench_r2f_lhbdc.err
python - <<PY
import json
d=json.load(open('gpurun_out/bench_r2f_lhbdc.json'))
print(d['value'], d['e2e']['value'], d['conv_tf32']['value'], d['parity'])
PY
( time timeout 600 python bench.py --impl reference --steps 3 --warmup 3 > gpurun_out/bench_r2f_ref.json 2> gpurun_out/bench_r2f_ref.err ) 2>&1 | tail -n 4; head -c 600 gpurun_out/bench_r2f_ref.json
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv --log-file gpurun_out/launches_r02.csv python bench.py --steps 1 --warmup 3 --no-e2e --no-cpu-baseline --no-tf32-leg --ncu-range > gpurun_out/bench_under_ncu.json 2> gpurun_out/bench_under_ncu.err; wc -l gpurun_out/launches_r02.csv
