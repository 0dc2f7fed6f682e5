#!/bin/bash
# round-2 GPU check #7: backward incl. full-size weight gradients; wgrad kernel timing + ncu; final bench lines
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_model.py -q -s -k 'phong or demo or golden' > gpurun_out/r02_run7_phong.log 2>&1; echo "phong rc=$?"
grep -E "passed|failed|^E  |Error" gpurun_out/r02_run7_phong.log | head -20
timeout 1500 python -m pytest tests/test_gpu_backward.py -q -s > gpurun_out/r02_run7_backward.log 2>&1; echo "backward rc=$?"
grep -E "passed|failed|dL/d|^E  |Error" gpurun_out/r02_run7_backward.log | head -40
timeout 600 python scripts/wgrad_time.py 2>&1 | tee gpurun_out/r02_wgrad_time.log
timeout 600 ncu --set full --clock-control none --import-source on -k regex:wgrad2d_kernel -s 3 -c 1 -f -o gpurun_out/r02_wgrad_trunk python scripts/wgrad_time.py > /dev/null 2>&1; echo "ncu wgrad rc=$?"
for c in 2 4 5; do
timeout 900 python bench.py --steps 10 --config $c > gpurun_out/r02_run7_bench_c$c.json 2> gpurun_out/r02_run7_bench_c$c.err; echo "bench c$c rc=$?"
done
python - <<'PY'
import json
for c in (2,4,5):
    d=json.load(open(f"gpurun_out/r02_run7_bench_c{c}.json")); print(f"c{c}", d["metric"], round(d["value"],1), "ms", round(d["ms_per_step"],2), "e2e", round(d["e2e"]["value"],1), "fast", round(d["other_precision"]["value"],1), "roof", round(d["roofline"]["frac"],3), "cpu", d["cpu_baseline"]["value"])
PY
