#!/bin/bash
# round-2 multi-GPU run B on ONE 8-GPU box: non-overlapped NCCL gather (nccl_sync, the new default) vs overlapped with few NCCL CTAs
mkdir -p gpurun_out
N=${1:-8}
run() {  # tag, env, extra args
  tag=$1; envs=$2; shift; shift
  env $envs timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29533 \
     bench.py --gpus $N --steps 10 --warmup 3 --no-other-precision --no-cpu-baseline "$@" > gpurun_out/r02_scale${N}b_$tag.json 2> gpurun_out/r02_scale${N}b_$tag.err
  echo "$tag rc=$?"; python - <<PY
import json
try:
    d=json.load(open("gpurun_out/r02_scale${N}b_$tag.json"))
    print("  value", round(d["value"],1), "ms/step", round(d["ms_per_step"],2), "per-rank", [round(x,2) for x in d["per_rank_ms"]], "e2e", round(d["e2e"]["value"],1), d["config"]["parallelism"][:70])
except Exception as e: print("  ERR", e)
PY
}
run fast_sync "A=1" --precision fast --gather nccl_sync
run fast_overlap_ctas2 "NCCL_MAX_CTAS=2" --precision fast --gather nccl
run exact_sync "A=1" --precision exact --gather nccl_sync
run fast_none "A=1" --precision fast --gather none
timeout 600 python bench.py --steps 10 --no-cpu-baseline > gpurun_out/r02_scale${N}b_n1_samebox.json 2> gpurun_out/r02_scale${N}b_n1_samebox.err; echo "n1 rc=$?"
python - <<PY
import json
d=json.load(open("gpurun_out/r02_scale${N}b_n1_samebox.json")); print("N=1 same box exact", round(d["value"],1), "fast", round(d["other_precision"]["value"],1))
PY
