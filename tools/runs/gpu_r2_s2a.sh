#!/bin/bash
# round 2, session 2, one GPU: the watchdog fix under stress (1/8 and 1/16 shards, deep pipelines), the GPU suite, the probe alone
mkdir -p gpurun_out
OUT=gpurun_out/r2s2_wd.log; : > $OUT
timeout 300 python tools/quick_shard.py --ranks 8 --depth 32 --frames 2000 --reps 8 2>&1 | tail -9 | cut -c1-900 >> $OUT
timeout 300 python tools/quick_shard.py --ranks 8 --depth 20 --frames 2000 --reps 6 2>&1 | tail -7 | cut -c1-900 >> $OUT
timeout 300 python tools/quick_shard.py --ranks 16 --depth 32 --frames 2000 --reps 4 2>&1 | tail -5 | cut -c1-900 >> $OUT
cat $OUT
timeout 1500 python -m pytest tests -m gpu -x -q --timeout 280 2>&1 | tail -8 > gpurun_out/r2s2_pytest.log; tail -4 gpurun_out/r2s2_pytest.log
timeout 200 python tools/scale_probe.py --depth 4,8 --frames 300 > gpurun_out/r2s2_probe_n1.log 2>&1; tail -5 gpurun_out/r2s2_probe_n1.log
