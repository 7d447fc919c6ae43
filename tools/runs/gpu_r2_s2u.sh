#!/bin/bash
# round 2, session 2, two GPUs: the bench lines with the mean of windows, one GPU (full default line) and two
mkdir -p gpurun_out
CUDA_VISIBLE_DEVICES=0 timeout 600 python bench.py --watchdog 500 > gpurun_out/r2s2_bench_n1_mean.log 2>&1
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29532 bench.py --gpus 2 --steps 20 --warmup 5 --no-extra --watchdog 240 > gpurun_out/r2s2_bench_n2_mean.log 2>&1
for f in n1_mean n2_mean; do python - <<PY
import json
try:
    j=json.loads([l for l in open("gpurun_out/r2s2_bench_$f.log").read().strip().splitlines() if l.startswith("{")][-1])
    w=j["config"]["windows_ms_per_step"]
    print("$f", round(j["value"]), "Mrays/s", round(j["ms_per_step"],4), "ms/step", "inflight", j["config"]["frames_in_flight"], "e2e", round(j["e2e"]["value"]), "e2e8", round(j.get("e2e_rgba8",{}).get("value",0)), {k: round(v,4) for k,v in w.items() if k!="in_order"}, "equal1gpu", j.get("frame_equal_to_1gpu"), "frac", round(j["roofline"]["frac"],3), round(j["roofline"]["frac_pipelined"],3), "traffic", j["roofline"]["traffic"], j["roofline"]["whole_frame_dram_bytes"], "cpu", (j.get("cpu_baseline") or {}).get("value"))
except Exception as e:
    print("$f failed", e); print(open("gpurun_out/r2s2_bench_$f.log").read()[-2000:])
PY
done
