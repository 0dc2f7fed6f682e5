#!/bin/bash
# round-2 GPU check #5: backward tests, config-4 input fusion (tests + bench + launch list)
mkdir -p gpurun_out
timeout 1500 python -m pytest tests/test_gpu_backward.py -q -s > gpurun_out/r02_run5_backward.log 2>&1; echo "backward rc=$?"
grep -E "passed|failed|err |cosine|^E  |Error" gpurun_out/r02_run5_backward.log | head -40
timeout 900 python -m pytest tests/test_gpu_model.py -q -s -k "texture or pretrained" > gpurun_out/r02_run5_texture.log 2>&1; echo "texture tests rc=$?"
grep -E "passed|failed|^E  |Error" gpurun_out/r02_run5_texture.log | head
timeout 900 python bench.py --steps 10 --config 4 > gpurun_out/r02_run5_bench_c4.json 2> gpurun_out/r02_run5_bench_c4.err; echo "bench c4 rc=$?"; tail -3 gpurun_out/r02_run5_bench_c4.err
python - <<'PY'
import json
d=json.load(open("gpurun_out/r02_run5_bench_c4.json")); print("c4", round(d["value"],1), "ms", round(d["ms_per_step"],2), "e2e", round(d["e2e"]["value"],1), "fast", round(d["other_precision"]["value"],1), "launches", d["gpu_launches_per_step"])
PY
cat > /tmp/prof_tex.py <<'PY'
import sys, numpy as np, torch
sys.path.insert(0, ".")
from rendernet_b200.engine import TextureRenderEngine
prec = sys.argv[1]
B = 24
eng = TextureRenderEngine(None, B, use_graph=False, seed=0, precision=prec)
rng = np.random.default_rng(0)
eng.upload((rng.random((B,64,64,64,1))<0.1).astype(np.float32), rng.standard_normal((B,199)).astype(np.float32),
           np.stack([rng.uniform(0,6.28,B), rng.uniform(-1,1,B), rng.uniform(0.8,1.3,B)],1).astype(np.float32))
eng.step_device(); torch.cuda.synchronize()
torch.cuda.profiler.start(); eng.step_device(); torch.cuda.synchronize(); torch.cuda.profiler.stop()
PY
for prec in fast exact; do
timeout 600 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r02_launches_texture_B24_$prec.csv python /tmp/prof_tex.py $prec > /dev/null 2>&1; echo "ncu tex $prec rc=$?"
done
python - <<'PY'
import csv,re
for prec in ("fast","exact"):
    lines=[l for l in open(f"gpurun_out/r02_launches_texture_B24_{prec}.csv") if not l.startswith("==")]
    rows=[r for r in csv.DictReader(lines) if r.get("Metric Name")=="gpu__time_duration.sum"]
    ms=[(re.sub(r"\(.*","",r["Kernel Name"])[:50], float(r["Metric Value"].replace(",",""))/(1e6 if r["Metric Unit"]=="ns" else 1e3 if r["Metric Unit"]=="us" else 1)) for r in rows]
    print(prec, "launches", len(ms), "total", round(sum(m for _,m in ms),2)); print("  first 8:", [(n,round(m,3)) for n,m in ms[:8]])
PY
