#!/bin/bash
# bit-identity of the totals across world sizes: the same 65-frame sequence (8 GOP-8) on W ranks, cudnn.benchmark off
cd "$(dirname "$0")/.."
W=$1
if [ "$W" = "1" ]; then
  python bench.py --sequence-frames 65 --deterministic --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/bench_r2_det65_1gpu.json 2> gpurun_out/bench_r2_det65_1gpu.err
else
  python -m torch.distributed.run --nnodes=1 --nproc-per-node $W --master-addr 127.0.0.1 --master-port 29633 bench.py --gpus $W --sequence-frames 65 --deterministic --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/bench_r2_det65_${W}gpu.json 2> gpurun_out/bench_r2_det65_${W}gpu.err
fi
python - <<PY
import json
p=json.load(open("gpurun_out/bench_r2_det65_${W}gpu.json"))
print("W=$W", p["value"], p["quality"])
PY
tail -n 2 gpurun_out/bench_r2_det65_${W}gpu.err
