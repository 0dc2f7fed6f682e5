#!/bin/bash
# 8-GPU run C: where do the ~3.2 ms per step go?  per-step device timestamps (compute vs gather) per rank
mkdir -p gpurun_out
N=${1:-8}
for g in nccl_sync nccl none; do
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29533 \
   bench.py --gpus $N --steps 10 --warmup 3 --no-other-precision --no-cpu-baseline --precision fast --gather $g --phases > gpurun_out/r02_scale${N}c_fast_$g.json 2> gpurun_out/r02_scale${N}c_fast_$g.err
echo "$g rc=$?"; python - <<PY
import json
d=json.load(open("gpurun_out/r02_scale${N}c_fast_$g.json")); print("  value", round(d["value"],1), "ms/step", round(d["ms_per_step"],2), "phases", d["phases"])
PY
done
