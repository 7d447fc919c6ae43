#!/bin/bash
mkdir -p gpurun_out
OUT=gpurun_out/r2_tune5.log; : > $OUT
for pol in 0 1 2; do for keep in 8 37; do PGRT_POOL_POLICY=$pol PGRT_KEEP_CTAS=$keep timeout 120 python tools/quick_c2.py --depth 4 >> $OUT 2>&1; done; done
for pol in 0 1 2; do PGRT_POOL_POLICY=$pol timeout 120 python tools/prof_frame.py --workload c2 --frames 3 2>&1 | grep "^level" >> $OUT; done
PGRT_POOL_POLICY=0 PGRT_KEEP_CTAS=37 timeout 120 python tools/quick_c2.py --depth 8 >> $OUT 2>&1
cat $OUT
