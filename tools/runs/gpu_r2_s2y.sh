#!/bin/bash
# round 2, session 2, one GPU: last knobs under the hybrid on the whole C2 frame: grid of the pool-only k_frame, k_trace CTAs per SM, pipeline depth
mkdir -p gpurun_out
OUT=gpurun_out/r2s2_last_knobs.log; : > $OUT
run() { env $1 timeout 100 python tools/scale_probe.py --depth $2 --frames 500 --variants solo 2>&1 | grep -E "N=|rror" | sed "s/^/$1 /" >> $OUT; }
run X=1 16
run PGRT_FRAME_CTAS_PER_SM=2 16
run PGRT_FRAME_CTAS_PER_SM=1 16
run PGRT_TRACE_CTAS_PER_SM=5 16
run PGRT_TRACE_CTAS_PER_SM=8 16
run X=1 32
cat $OUT
