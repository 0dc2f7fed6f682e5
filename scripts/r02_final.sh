#!/bin/bash
# round-2 final validation on a fresh box: the whole GPU suite, smoke(), the default bench line, configs 4 / 5, the reference arm
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q > gpurun_out/r02_final_gpu_suite.log 2>&1; echo "gpu suite rc=$?"
tail -4 gpurun_out/r02_final_gpu_suite.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -2
timeout 900 python bench.py > gpurun_out/r02_final_bench_c2.json 2> gpurun_out/r02_final_bench_c2.err; echo "bench rc=$?"
for c in 4 5; do
timeout 600 python bench.py --config $c --steps 10 > gpurun_out/r02_final_bench_c$c.json 2> gpurun_out/r02_final_bench_c$c.err; echo "bench c$c rc=$?"
done
python - <<'PY'
import json
for c in (2,4,5):
    try:
        d=json.load(open(f"gpurun_out/r02_final_bench_c{c}.json")); print(f"c{c}", d["metric"], round(d["value"],1), "ms", round(d["ms_per_step"],2), "e2e", round(d["e2e"]["value"],1), "fast", round(d["other_precision"]["value"],1), "roof", round(d["roofline"]["frac"],3), "cpu", d["cpu_baseline"]["value"], "launches", d["gpu_launches"])
    except Exception as e:
        print("c", c, "failed", e)
PY
