#!/bin/bash
mkdir -p gpurun_out
OUT=gpurun_out/r2_shard2.log; : > $OUT
for d in 16 24 32; do timeout 120 python tools/quick_shard.py --ranks 8 --depth $d >> $OUT 2>&1; done
for d in 16 32; do timeout 120 python tools/quick_shard.py --ranks 4 --depth $d >> $OUT 2>&1; done
for d in 8 16 32; do timeout 120 python tools/quick_shard.py --ranks 2 --depth $d >> $OUT 2>&1; done
for d in 4 8 16; do timeout 120 python tools/quick_shard.py --ranks 1 --depth $d >> $OUT 2>&1; done
PGRT_POOL_POLICY=2 timeout 120 python tools/quick_shard.py --ranks 8 --depth 16 --tag policy2 >> $OUT 2>&1
PGRT_POOL_POLICY=2 timeout 120 python tools/quick_shard.py --ranks 8 --depth 32 --tag policy2 >> $OUT 2>&1
cat $OUT
