#!/bin/bash
# round-2 GPU check #9: where the training step's time goes (ncu launch list) + ncu --set full of the two weight-gradient kernels
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_model.py -q -k phong 2>&1 | tail -3
timeout 600 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r02_launches_train_step_exact.csv python scripts/profile_train_step.py --precision exact > /dev/null 2>&1; echo "launch list rc=$?"
timeout 400 ncu --profile-from-start off --set full --clock-control none --import-source on -k regex:wgrad2d_kernel -s 40 -c 1 -f -o gpurun_out/r02_wgrad2d_train python scripts/profile_train_step.py --precision exact > /dev/null 2>&1; echo "ncu wgrad2d rc=$?"
timeout 400 ncu --profile-from-start off --set full --clock-control none --import-source on -k regex:wgrad_direct_kernel -s 2 -c 1 -f -o gpurun_out/r02_wgrad_direct_train python scripts/profile_train_step.py --precision exact > /dev/null 2>&1; echo "ncu wgrad_direct rc=$?"
python - <<'PY'
import csv, collections
rows=[r for r in csv.reader(open("gpurun_out/r02_launches_train_step_exact.csv")) if len(r)>10]
hdr=rows[0]; ki=hdr.index("Kernel Name"); vi=hdr.index("Metric Value")
agg=collections.defaultdict(lambda:[0,0.0])
for r in rows[1:]:
    try: v=float(r[vi].replace(",",""))
    except ValueError: continue
    n=r[ki].split("(")[0][:60]; agg[n][0]+=1; agg[n][1]+=v
tot=sum(v for _,v in agg.values())
print("launches", sum(c for c,_ in agg.values()), "total ms", round(tot/1e6,2))
for n,(c,v) in sorted(agg.items(), key=lambda kv:-kv[1][1])[:14]:
    print(f"{v/1e6:8.2f} ms {100*v/tot:5.1f}% x{c:4d}  {n}")
PY
