#!/bin/bash
mkdir -p gpurun_out
OUT=gpurun_out/r2_wd.log; : > $OUT
for i in 1 2 3; do for d in 20 24 32; do timeout 200 python tools/quick_shard.py --ranks 8 --depth $d --frames 800 2>&1 | tail -1 | cut -c1-900 >> $OUT; done; done
timeout 200 python tools/quick_shard.py --ranks 4 --depth 32 --frames 800 2>&1 | tail -1 | cut -c1-900 >> $OUT
timeout 200 python tools/quick_shard.py --ranks 16 --depth 32 --frames 800 2>&1 | tail -1 | cut -c1-900 >> $OUT
cat $OUT
