#!/bin/bash
# 8-GPU run D: copy-engine peer gather with one stream per destination (no SMs), overlapped with the next step
mkdir -p gpurun_out
N=${1:-8}
for spec in "fast peer" "exact peer"; do
set -- $spec
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29533 \
   bench.py --gpus $N --steps 10 --warmup 3 --no-other-precision --no-cpu-baseline --precision $1 --gather $2 --phases > gpurun_out/r02_scale${N}d_$1_$2.json 2> gpurun_out/r02_scale${N}d_$1_$2.err
echo "$1 $2 rc=$?"; tail -2 gpurun_out/r02_scale${N}d_$1_$2.err; python - <<PY
import json
d=json.load(open("gpurun_out/r02_scale${N}d_$1_$2.json")); print("  value", round(d["value"],1), "ms/step", round(d["ms_per_step"],2), d["config"]["parallelism"][:80], "phases", d["phases"])
PY
done
