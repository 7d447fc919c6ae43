#!/bin/bash
# round 2, session 2, N GPUs ($1): the bench line as the driver runs it
N=$1; mkdir -p gpurun_out
S=$(date +%s)
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 2955$N bench.py --gpus $N --steps 20 --warmup 5 --watchdog 500 > gpurun_out/r2s2_bench_n$N.log 2>&1
echo "bench N=$N took $(( $(date +%s) - S )) s"
python - <<PY
import json
try:
    j=json.loads([l for l in open("gpurun_out/r2s2_bench_n$N.log").read().strip().splitlines() if l.startswith("{")][-1])
    print("N=$N", round(j["value"]), "Mrays/s", round(j["ms_per_step"],4), "ms/step", "inflight", j["config"]["frames_in_flight"], "e2e", round(j["e2e"]["value"]), "e2e8", round(j.get("e2e_rgba8",{}).get("value",0)), j["config"]["completion"], j["config"]["windows_ms_per_step"], "equal1gpu", j.get("frame_equal_to_1gpu"), "host_us", round(j["config"]["host_issue_us_per_step"],1), j["config"]["host_issue_parts_us"], j["roofline"]["scheduler"])
    print(json.dumps(j["config"].get("extra"))[:1500])
except Exception as e:
    print("failed", e); print(open("gpurun_out/r2s2_bench_n$N.log").read()[-2500:])
PY
