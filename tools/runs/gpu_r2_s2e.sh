#!/bin/bash
# round 2, session 2, eight GPUs: the probe and the bench line as the driver runs it
mkdir -p gpurun_out
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29548 tools/scale_probe.py --depth 16,32 --frames 800 --variants counter,solo > gpurun_out/r2s2_probe_n8.log 2>&1
grep -E "^N=|rror" gpurun_out/r2s2_probe_n8.log | tail
S=$(date +%s)
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29549 bench.py --gpus 8 --steps 20 --warmup 5 --watchdog 500 > gpurun_out/r2s2_bench_n8.log 2>&1
echo "bench N=8 took $(( $(date +%s) - S )) s"
python - <<PY
import json
try:
    j=json.loads([l for l in open("gpurun_out/r2s2_bench_n8.log").read().strip().splitlines() if l.startswith("{")][-1])
    print("N=8", round(j["value"]), "Mrays/s", round(j["ms_per_step"],4), "ms/step", "inflight", j["config"]["frames_in_flight"], "e2e", round(j["e2e"]["value"]), "e2e8", round(j.get("e2e_rgba8",{}).get("value",0)), j["config"]["completion"], j["config"]["windows_ms_per_step"], "equal1gpu", j.get("frame_equal_to_1gpu"), "host_us", round(j["config"]["host_issue_us_per_step"],1), j["config"]["host_issue_parts_us"])
    print(json.dumps(j["config"].get("extra"))[:1500])
except Exception as e:
    print("failed", e); print(open("gpurun_out/r2s2_bench_n8.log").read()[-2500:])
PY
