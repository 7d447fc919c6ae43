#!/bin/bash
# round 2, session 2, one GPU: hybrid with direct pixel stores (no colour queue, no k_resolve at one sample per pixel); 2 keeper CTAs
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q --timeout 280 2>&1 | tail -4
OUT=gpurun_out/r2s2_direct.log; : > $OUT
timeout 200 python tools/scale_probe.py --depth 16 --frames 600 --variants solo 2>&1 | grep -E "N=|rror" >> $OUT
for r in 2 4; do timeout 200 python tools/quick_shard.py --ranks $r --depth 16 --frames 1200 2>&1 | tail -1 | cut -c1-300 >> $OUT; done
timeout 200 python tools/quick_shard.py --ranks 8 --depth 32 --frames 1500 2>&1 | tail -1 | cut -c1-300 >> $OUT
cat $OUT
timeout 900 python bench.py --watchdog 800 --no-extra > gpurun_out/r2s2_bench_direct.log 2>&1
python - <<PY
import json
try:
    j=json.loads([l for l in open("gpurun_out/r2s2_bench_direct.log").read().strip().splitlines() if l.startswith("{")][-1])
    print("N=1", round(j["value"]), "Mrays/s", round(j["ms_per_step"],4), "ms/step", "inflight", j["config"]["frames_in_flight"], "e2e", round(j["e2e"]["value"]), "e2e8", round(j.get("e2e_rgba8",{}).get("value",0)), j["config"]["windows_ms_per_step"], "unpip", j["roofline"]["frame_ms_unpipelined"])
except Exception as e:
    print("failed", e); print(open("gpurun_out/r2s2_bench_direct.log").read()[-3000:])
PY
