#!/bin/bash
# round 2, session 2, one GPU: where the warps of k_frame spend their time (build with -DPGRT_FRAME_TIMING); the wavefront scheduler for comparison
mkdir -p gpurun_out
OUT=gpurun_out/r2s2_timing.log; : > $OUT
for d in 1 8; do PGRT_LIB=$PWD/build/variants/timing.so timeout 120 python tools/quick_c2.py --depth $d --tag timing_depth$d >> $OUT 2>&1; done
PGRT_LIB=$PWD/build/variants/timing.so timeout 120 python tools/quick_c2.py --depth 8 --tag timing_lambert --params '{"shader_mode": 1}' >> $OUT 2>&1
PGRT_LIB=$PWD/build/variants/timing.so timeout 120 python tools/quick_c2.py --depth 8 --tag timing_maxdepth1 --params '{"max_depth": 1}' >> $OUT 2>&1
PGRT_KEEP_CTAS=1 PGRT_LIB=$PWD/build/variants/timing.so timeout 120 python tools/quick_c2.py --depth 8 --tag timing_keep1 >> $OUT 2>&1
for d in 8 16; do timeout 120 python tools/quick_c2.py --depth $d --tag wavefront_depth$d --params '{"scheduler": 1}' >> $OUT 2>&1; done
cut -c1-420 $OUT
