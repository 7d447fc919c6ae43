#!/bin/bash
# round 2: A/B of k_frame builds (register caps, cold paths inlined or not) and the cost of the secondary rays by depth
mkdir -p gpurun_out
OUT=gpurun_out/r2_tune.log; : > $OUT
for lib in build/variants/*.so; do PGRT_LIB=$PWD/$lib timeout 120 python tools/quick_c2.py --tag $(basename $lib .so) >> $OUT 2>&1; done
for d in 0 1 2 4; do timeout 120 python tools/quick_c2.py --tag maxdepth$d --params "{\"max_depth\": $d}" >> $OUT 2>&1; done
timeout 120 python tools/quick_c2.py --tag lambert --params '{"shader_mode": 1}' >> $OUT 2>&1
for pat in 1 16; do PGRT_CLAIM_PATIENCE=$pat timeout 120 python tools/quick_c2.py --tag patience$pat >> $OUT 2>&1; done
cat $OUT
timeout 600 python -m pytest tests -m gpu -x -q --timeout 240 -k "schedulers or overflow or pipelined or golden or C2_full" 2>&1 | tail -3
