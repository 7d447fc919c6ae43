#!/bin/bash
# round 2: the whole GPU suite, then the default bench line with the C3 / C5 extras
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q --timeout 280 2>&1 | tail -8 > gpurun_out/r2_pytest.log; tail -4 gpurun_out/r2_pytest.log
timeout 900 python bench.py --watchdog 800 > gpurun_out/r2_bench_default.log 2>&1
python - <<PY
import json
try:
    j=json.loads([l for l in open("gpurun_out/r2_bench_default.log").read().strip().splitlines() if l.startswith("{")][-1])
    print("N=1", round(j["value"]), "Mrays/s", round(j["ms_per_step"],4), "ms/step", "e2e", round(j["e2e"]["value"]), "e2e8", round(j.get("e2e_rgba8",{}).get("value",0)), j["config"]["windows_ms_per_step"], "cpu", j.get("cpu_baseline"))
    print(json.dumps(j["config"].get("extra"), indent=1))
    print(json.dumps(j["roofline"], indent=1)[:1500])
except Exception as e:
    print("failed", e); print(open("gpurun_out/r2_bench_default.log").read()[-3000:])
PY
timeout 300 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/r2_bench_reference.log 2>&1; tail -c 600 gpurun_out/r2_bench_reference.log
