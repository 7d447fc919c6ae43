#!/bin/bash
mkdir -p gpurun_out
OUT=gpurun_out/r2_tune4.log; : > $OUT
timeout 120 python tools/prof_frame.py --workload c2 --frames 4 >> $OUT 2>&1
for d in 4 8; do timeout 120 python tools/quick_c2.py --tag ticket_primary --depth $d >> $OUT 2>&1; done
for pat in 1 16; do PGRT_CLAIM_PATIENCE=$pat timeout 120 python tools/quick_c2.py --tag patience$pat >> $OUT 2>&1; done
for c in 8 16; do PGRT_MIN_CLAIM=$c timeout 120 python tools/quick_c2.py --tag minclaim$c >> $OUT 2>&1; done
cat $OUT
timeout 600 python -m pytest tests -m gpu -x -q --timeout 240 -k "schedulers or overflow or pipelined or golden or C2_full or sharded" 2>&1 | tail -3
