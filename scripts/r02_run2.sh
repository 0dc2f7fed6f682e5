#!/bin/bash
# round-2 GPU check #2: all GPU tests, bench for configs 2/4/5 + reference arm, ncu launch list + full capture (exact mode)
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q -s > gpurun_out/r02_run2_gpu_suite.log 2>&1; echo "gpu suite rc=$?"
grep -E "passed|failed" gpurun_out/r02_run2_gpu_suite.log | tail -3; grep -E "^FAILED|IMAGE|config [245]|K=9216" gpurun_out/r02_run2_gpu_suite.log | head -40
timeout 900 python bench.py --steps 10 > gpurun_out/r02_run2_bench_c2.json 2> gpurun_out/r02_run2_bench_c2.err; echo "bench c2 rc=$?"
timeout 900 python bench.py --steps 10 --config 4 > gpurun_out/r02_run2_bench_c4.json 2> gpurun_out/r02_run2_bench_c4.err; echo "bench c4 rc=$?"; tail -3 gpurun_out/r02_run2_bench_c4.err
timeout 900 python bench.py --steps 3 --config 5 > gpurun_out/r02_run2_bench_c5.json 2> gpurun_out/r02_run2_bench_c5.err; echo "bench c5 rc=$?"; tail -3 gpurun_out/r02_run2_bench_c5.err
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/r02_run2_bench_ref.json 2> gpurun_out/r02_run2_bench_ref.err; echo "bench ref rc=$?"
python - <<'PY'
import json
for c in ("c2","c4","c5","ref"):
    try:
        d=json.load(open(f"gpurun_out/r02_run2_bench_{c}.json"))
        print(c, d["metric"], round(d["value"],2), "ms/step", round(d["ms_per_step"],2), "e2e", round(d["e2e"]["value"],2),
              "other", (d.get("other_precision") or {}).get("value"), "roof", (d.get("roofline") or {}).get("frac"),
              "cpu", (d.get("cpu_baseline") or {}).get("value"), (d.get("cpu_baseline") or {}).get("cores"))
    except Exception as e: print(c, "ERR", e)
PY
timeout 900 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r02_launches_step_B24_exact.csv python scripts/profile_step.py --batch 24 --precision exact > gpurun_out/r02_run2_ncu1.log 2>&1; echo "ncu launches rc=$?"
timeout 900 ncu --profile-from-start off --set full --clock-control none --import-source on -k regex:igemm_kernel -s 23 -c 2 -f -o gpurun_out/r02_exact_proj_trunk python scripts/profile_step.py --batch 24 --precision exact > gpurun_out/r02_run2_ncu2.log 2>&1; echo "ncu full rc=$?"
ls -la gpurun_out | tail -20
