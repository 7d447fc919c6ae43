#!/bin/bash
# round 2, two GPUs: the sharded paths against the single-GPU frame, then a short scaling bench
mkdir -p gpurun_out
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29531 tools/p2p_check.py > gpurun_out/r2_p2p_n2.log 2>&1
grep -E "ranks|Error|error|assert" gpurun_out/r2_p2p_n2.log | tail -12
timeout 400 python bench.py --gpus 1 --steps 20 --warmup 5 --min-seconds 0.5 --no-extra --no-cpu-baseline > gpurun_out/r2_bench_n1.log 2>&1
timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29532 bench.py --gpus 2 --steps 20 --warmup 5 --min-seconds 0.5 --no-extra > gpurun_out/r2_bench_n2.log 2>&1
for n in 1 2; do python - <<PY
import json
try:
    j=json.loads([l for l in open("gpurun_out/r2_bench_n$n.log").read().strip().splitlines() if l.startswith("{")][-1])
    print("N=$n", round(j["value"]), "Mrays/s", round(j["ms_per_step"],4), "ms/step", "e2e", round(j["e2e"]["value"]), "e2e8", round(j.get("e2e_rgba8",{}).get("value",0)), j["config"]["completion"], j["config"]["windows_ms_per_step"], "equal1gpu", j.get("frame_equal_to_1gpu"), "host_us", round(j["config"]["host_issue_us_per_step"],1))
except Exception as e:
    print("N=$n failed", e); print(open("gpurun_out/r2_bench_n$n.log").read()[-3000:])
PY
done
