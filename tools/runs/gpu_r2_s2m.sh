#!/bin/bash
# round 2, session 2, one GPU: the suite, the default bench line (timed), the reference arm, the launch list of the bench command
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q --timeout 280 2>&1 | tail -4
S=$(date +%s)
timeout 900 python bench.py --watchdog 800 > gpurun_out/r2s2_bench_default.log 2>&1
echo "default bench took $(( $(date +%s) - S )) s"
python - <<PY
import json
try:
    j=json.loads([l for l in open("gpurun_out/r2s2_bench_default.log").read().strip().splitlines() if l.startswith("{")][-1])
    print("N=1", round(j["value"]), "Mrays/s", round(j["ms_per_step"],4), "ms/step", "inflight", j["config"]["frames_in_flight"], "e2e", round(j["e2e"]["value"]), "e2e8", round(j.get("e2e_rgba8",{}).get("value",0)), j["config"]["windows_ms_per_step"], "cpu", j.get("cpu_baseline"))
    print(json.dumps(j["config"].get("extra"))[:1200])
    print(json.dumps(j["roofline"])[:1800])
    print("launches", j["gpu_launches"], j["clocks"])
except Exception as e:
    print("failed", e); print(open("gpurun_out/r2s2_bench_default.log").read()[-3000:])
PY
timeout 300 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/r2s2_bench_reference.log 2>&1; tail -c 400 gpurun_out/r2s2_bench_reference.log
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 1500 --csv --log-file gpurun_out/r2s2_launches_bench.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-extra --min-seconds 0.02 > gpurun_out/r2s2_ncu_launches.log 2>&1
echo "ncu rc=$?"; wc -l gpurun_out/r2s2_launches_bench.csv
