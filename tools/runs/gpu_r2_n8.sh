#!/bin/bash
# round 2, eight GPUs: the scaling bench at N = 8 (two pipeline depths) and N = 4, flags, watchdog
mkdir -p gpurun_out
run() { # n inflight tag
  timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $1 --master-addr 127.0.0.1 --master-port 2954$1 bench.py --gpus $1 --steps 20 --warmup 5 --no-extra --watchdog 240 --inflight $2 > gpurun_out/r2_bench_$3.log 2>&1
  python - <<PY
import json
try:
    j=json.loads([l for l in open("gpurun_out/r2_bench_$3.log").read().strip().splitlines() if l.startswith("{")][-1])
    print("$3", round(j["value"]), "Mrays/s", round(j["ms_per_step"],4), "ms/step", "inflight", j["config"]["frames_in_flight"], "e2e", round(j["e2e"]["value"]), "e2e8", round(j.get("e2e_rgba8",{}).get("value",0)), j["config"]["completion"], j["config"]["windows_ms_per_step"], "equal1gpu", j.get("frame_equal_to_1gpu"), "host_us", round(j["config"]["host_issue_us_per_step"],1), j["config"]["host_issue_parts_us"])
except Exception as e:
    print("$3 failed", e); print(open("gpurun_out/r2_bench_$3.log").read()[-2500:])
PY
}
run 8 0 n8_auto
run 8 8 n8_d8
run 4 0 n4_auto
run 2 0 n2_auto
run 1 0 n1_auto
