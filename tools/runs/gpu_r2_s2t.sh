#!/bin/bash
# round 2, session 2, one GPU: the distribution of the bench's timed windows (median against mean), twice
mkdir -p gpurun_out
for i in 1 2; do
timeout 600 python bench.py --no-extra --no-cpu-baseline --watchdog 500 > gpurun_out/r2s2_bench_win$i.log 2>&1
python - <<PY
import json
j=json.loads([l for l in open("gpurun_out/r2s2_bench_win$i.log").read().strip().splitlines() if l.startswith("{")][-1])
w=j["config"]["windows_ms_per_step"]
print("run $i value", round(j["value"]), {k: round(v,4) for k,v in w.items() if k!="in_order"})
print(" ".join(f"{x:.3f}" for x in w["in_order"]))
PY
done
