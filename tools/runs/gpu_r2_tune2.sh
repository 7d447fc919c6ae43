#!/bin/bash
mkdir -p gpurun_out
OUT=gpurun_out/r2_tune2.log; : > $OUT
for keep in 19 37 74 148 296; do for d in 4; do PGRT_KEEP_CTAS=$keep timeout 120 python tools/quick_c2.py --tag keepctas$keep --depth $d >> $OUT 2>&1; done; done
for d in 2 3 6; do PGRT_KEEP_CTAS=74 timeout 120 python tools/quick_c2.py --tag keepctas74 --depth $d >> $OUT 2>&1; done
for lib in build/variants/*.so; do PGRT_KEEP_CTAS=74 PGRT_LIB=$PWD/$lib timeout 120 python tools/quick_c2.py --tag $(basename $lib .so)_keep74 >> $OUT 2>&1; done
cat $OUT
