#!/bin/bash
# round 2, session 2, one GPU: CTA size of k_frame (a CTA holds its SM slot until its last warp is done), depth of recursion
mkdir -p gpurun_out
OUT=gpurun_out/r2s2_cta.log; : > $OUT
timeout 120 python tools/quick_c2.py --depth 8 --tag t128 >> $OUT 2>&1
for v in t64 t32 t32mb16; do PGRT_LIB=$PWD/build/variants/$v.so timeout 120 python tools/quick_c2.py --depth 8 --tag $v >> $OUT 2>&1; done
for k in 2 4 16; do PGRT_KEEP_CTAS=$k PGRT_LIB=$PWD/build/variants/t32.so timeout 120 python tools/quick_c2.py --depth 8 --tag t32keep$k >> $OUT 2>&1; done
PGRT_POOL_POLICY=2 PGRT_LIB=$PWD/build/variants/t32.so timeout 120 python tools/quick_c2.py --depth 8 --tag t32policy2 >> $OUT 2>&1
for d in 0 1 2 4; do timeout 120 python tools/quick_c2.py --depth 8 --tag t128maxdepth$d --params "{\"max_depth\": $d}" >> $OUT 2>&1; PGRT_LIB=$PWD/build/variants/t32.so timeout 120 python tools/quick_c2.py --depth 8 --tag t32maxdepth$d --params "{\"max_depth\": $d}" >> $OUT 2>&1; done
cut -c1-400 $OUT
