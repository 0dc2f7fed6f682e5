#!/bin/bash
# round-2 GPU check #3: half-N MMAs on the banded conv3d edge blocks -- correctness (whole suite) and per-launch times
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q -x > gpurun_out/r02_run3_gpu_suite.log 2>&1; echo "gpu suite rc=$?"
grep -E "passed|failed" gpurun_out/r02_run3_gpu_suite.log | tail -3; grep -E "^FAILED|^E  " gpurun_out/r02_run3_gpu_suite.log | head -20
for prec in fast exact; do
timeout 900 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r02_launches_step_B24_${prec}_halfN.csv python scripts/profile_step.py --batch 24 --precision $prec > /dev/null 2>&1; echo "ncu $prec rc=$?"
done
python - <<'PY'
import csv,re
for prec in ("fast","exact"):
    lines=[l for l in open(f"gpurun_out/r02_launches_step_B24_{prec}_halfN.csv") if not l.startswith("==")]
    rows=[r for r in csv.DictReader(lines) if r.get("Metric Name")=="gpu__time_duration.sum"]
    ms=[float(r["Metric Value"].replace(",",""))/ (1e6 if r["Metric Unit"]=="ns" else 1e3 if r["Metric Unit"]=="us" else 1) for r in rows]
    print(prec, "launches", len(ms), "total", round(sum(ms),2), "banded res1 (launch 3..23):", [round(x,3) for x in ms[3:24]][:6], "sum", round(sum(ms[1:24]),2))
PY
timeout 600 python scripts/ab_step.py "exact|exact|" "fast|fast|" 2>&1 | tail -8
