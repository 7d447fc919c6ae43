#!/bin/bash
mkdir -p gpurun_out
OUT=gpurun_out/r2_tune3.log; : > $OUT
PGRT_KEEP_CTAS=19 timeout 120 python tools/prof_frame.py --workload c2 --frames 4 >> $OUT 2>&1
PGRT_KEEP_CTAS=148 timeout 120 python tools/prof_frame.py --workload c2 --frames 4 >> $OUT 2>&1
for keep in 4 8 19; do for d in 4 8; do PGRT_KEEP_CTAS=$keep timeout 120 python tools/quick_c2.py --tag keepctas$keep --depth $d >> $OUT 2>&1; done; done
cat $OUT
