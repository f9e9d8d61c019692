#!/bin/bash
cd "$(dirname "$0")/.."
python bench.py --workload ojsp_search_4k --steps 3 --warmup 3 > gpurun_out/bench_r2_ojsp_1gpu.json 2> gpurun_out/bench_r2_ojsp_1gpu.err; tail -c 700 gpurun_out/bench_r2_ojsp_1gpu.json; tail -n 3 gpurun_out/bench_r2_ojsp_1gpu.err
python bench.py --workload lhbdc_frames --steps 3 --warmup 3 > gpurun_out/bench_r2_frames_1gpu.json 2> gpurun_out/bench_r2_frames_1gpu.err; tail -c 900 gpurun_out/bench_r2_frames_1gpu.json; tail -n 3 gpurun_out/bench_r2_frames_1gpu.err
python bench.py --workload flex_gop16_allq --steps 1 --warmup 3 > gpurun_out/bench_r2_flex_1gpu.json 2> gpurun_out/bench_r2_flex_1gpu.err; tail -c 700 gpurun_out/bench_r2_flex_1gpu.json; tail -n 3 gpurun_out/bench_r2_flex_1gpu.err
