#!/bin/bash
# round 2, session 2, one GPU: per-kernel times of the C5 (10 M triangles) build
mkdir -p gpurun_out
timeout 300 python tools/build_only.py --workload c5 --reps 3 > gpurun_out/r2s2_build_c5.log 2>&1; tail -3 gpurun_out/r2s2_build_c5.log | cut -c1-400
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r2s2_build_c5_launches.csv python tools/build_only.py --workload c5 --reps 1 > gpurun_out/r2s2_build_c5_ncu.log 2>&1
python - <<PY
import csv, collections
rows=[r for r in csv.reader(open("gpurun_out/r2s2_build_c5_launches.csv")) if len(r)>5]
h=rows[0]; ik=h.index("Kernel Name"); iv=h.index("Metric Value")
t=collections.defaultdict(lambda:[0,0.0])
for r in rows[1:]:
    try: v=float(r[iv].replace(",",""))
    except: continue
    k=r[ik].split("(")[0]; t[k][0]+=1; t[k][1]+=v
tot=sum(v[1] for v in t.values())
print("unit", rows[1][h.index("Metric Unit")], "total", tot)
for k,v in sorted(t.items(), key=lambda kv:-kv[1][1]): print(f"{k:40s} {v[0]:4d} launches {v[1]:12.1f}")
PY
