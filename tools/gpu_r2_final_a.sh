#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q 2>&1 | tail -n 15 | tee gpurun_out/pytest_r2e.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -n 5 | tee gpurun_out/smoke_r2e.log
timeout 900 python bench.py --steps 3 --warmup 3 > gpurun_out/bench_r2e_lhbdc.json 2> gpurun_out/bench_r2e_lhbdc.err; tail -n 3 gpurun_out/bench_r2e_lhbdc.err; head -c 1500 gpurun_out/bench_r2e_lhbdc.json
