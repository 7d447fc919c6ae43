#!/bin/bash
# round 2, session 2, one GPU: build with per-digit row scans; the L2 flush's cost and the persisting-L2 window at N = 1; GPU suite
mkdir -p gpurun_out
timeout 300 python tools/build_only.py --workload c5 --reps 3 > gpurun_out/r2s2_build_c5b.log 2>&1; tail -2 gpurun_out/r2s2_build_c5b.log | cut -c1-400
timeout 300 python tools/build_only.py --workload c4 --reps 3 2>&1 | tail -1 | cut -c1-400
timeout 300 python tools/build_only.py --workload c2 --reps 3 2>&1 | tail -1 | cut -c1-400
OUT=gpurun_out/r2s2_probe_n1b.log; : > $OUT
timeout 200 python tools/scale_probe.py --depth 4,8,16 --flush 1,0 --frames 400 --variants solo >> $OUT 2>&1
echo "== PGRT_L2_PERSIST=1" >> $OUT
PGRT_L2_PERSIST=1 timeout 200 python tools/scale_probe.py --depth 8 --flush 1,0 --frames 400 --variants solo >> $OUT 2>&1
grep -E "^N=|^==|rror" $OUT
timeout 1500 python -m pytest tests -m gpu -x -q --timeout 280 2>&1 | tail -4
