#!/bin/bash
# first GPU contact of round 2: GPU tests, per-level statistics of one C2 frame, a short bench line
set -x
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -15 > gpurun_out/r2_pytest.log
timeout 300 python tools/prof_frame.py --workload c2 --frames 5 --stats gpurun_out/r2_level_stats_c2.json > gpurun_out/r2_prof_c2.log 2>&1
for d in 1 2 4 8; do
  timeout 300 python bench.py --steps 100 --warmup 5 --inflight $d --no-cpu-baseline > gpurun_out/r2_bench_c2_d$d.log 2>&1
done
tail -3 gpurun_out/r2_pytest.log; cat gpurun_out/r2_prof_c2.log; for d in 1 2 4 8; do python - <<PY
import json
try:
    j=json.loads(open("gpurun_out/r2_bench_c2_d$d.log").read().strip().splitlines()[-1])
    print("inflight $d", round(j["value"]), j["ms_per_step"], "e2e", round(j["e2e"]["value"]), "unpip", j["roofline"]["frame_ms_unpipelined"])
except Exception as e:
    print("inflight $d failed", e); print(open("gpurun_out/r2_bench_c2_d$d.log").read()[-2000:])
PY
done
