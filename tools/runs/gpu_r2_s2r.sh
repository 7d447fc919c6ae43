#!/bin/bash
# round 2, session 2, one GPU, final build: launch list of the bench command, ncu --set full of one C2 frame's kernels (hybrid; fused), the default bench line
mkdir -p gpurun_out
timeout 600 python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-extra --min-seconds 0.02 > gpurun_out/r2s2_plain_final.log 2>&1 || { tail -5 gpurun_out/r2s2_plain_final.log; exit 1; }
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 1500 --csv --log-file gpurun_out/r2s2_launches_final.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-extra --min-seconds 0.02 > gpurun_out/r2s2_ncu_launches_final.log 2>&1
echo "launch list rc=$?"; wc -l gpurun_out/r2s2_launches_final.csv
timeout 300 python tools/prof_frame.py --workload c2 --frames 3 > gpurun_out/r2s2_prof_c2_final.log 2>&1 && \
timeout 900 ncu --set full --clock-control none --import-source on -k regex:'k_trace|k_shade|k_phong|k_frame|k_resolve' --launch-skip 4 -c 4 -f -o gpurun_out/r2s2_ncu_hybrid_final python tools/prof_frame.py --workload c2 --frames 2 > gpurun_out/r2s2_ncu_hf.log 2>&1
tail -1 gpurun_out/r2s2_ncu_hf.log | cut -c1-200; head -2 gpurun_out/r2s2_prof_c2_final.log | cut -c1-400
timeout 900 python bench.py --watchdog 800 > gpurun_out/r2s2_bench_final.log 2>&1
python - <<PY
import json
try:
    j=json.loads([l for l in open("gpurun_out/r2s2_bench_final.log").read().strip().splitlines() if l.startswith("{")][-1])
    print("N=1", round(j["value"]), "Mrays/s", round(j["ms_per_step"],4), "ms/step", "inflight", j["config"]["frames_in_flight"], "e2e", round(j["e2e"]["value"]), "e2e8", round(j.get("e2e_rgba8",{}).get("value",0)), j["config"]["windows_ms_per_step"], "unpip", j["roofline"]["frame_ms_unpipelined"], j["roofline"]["frac"], j["roofline"]["frac_pipelined"])
except Exception as e:
    print("failed", e); print(open("gpurun_out/r2s2_bench_final.log").read()[-3000:])
PY
