#!/bin/bash
# round-2 GPU check #6: wgrad (tcgen05, MN-major) + backward + texture decoder tiled conv; whole suite; bench config 4
mkdir -p gpurun_out
timeout 1500 python -m pytest tests/test_gpu_backward.py -q -s > gpurun_out/r02_run6_backward.log 2>&1; echo "backward rc=$?"
grep -E "passed|failed|err |cosine|^E  |Error" gpurun_out/r02_run6_backward.log | head -50
timeout 1500 python -m pytest tests -m gpu -q --deselect tests/test_gpu_backward.py > gpurun_out/r02_run6_gpu_suite.log 2>&1; echo "suite rc=$?"
grep -E "passed|failed|^FAILED" gpurun_out/r02_run6_gpu_suite.log | tail -8
timeout 900 python bench.py --steps 10 --config 4 > gpurun_out/r02_run6_bench_c4.json 2> gpurun_out/r02_run6_bench_c4.err; echo "bench c4 rc=$?"; tail -3 gpurun_out/r02_run6_bench_c4.err
python - <<'PY'
import json
d=json.load(open("gpurun_out/r02_run6_bench_c4.json")); print("c4", round(d["value"],1), "ms", round(d["ms_per_step"],2), "e2e", round(d["e2e"]["value"],1), "fast", round(d["other_precision"]["value"],1), "launches", d["gpu_launches_per_step"])
PY
