#!/bin/bash
# round 2, session 2, one GPU: hybrid on an eighth of the frame with smaller k_trace grids; flush grid under the hybrid on the whole frame
mkdir -p gpurun_out
OUT=gpurun_out/r2s2_small_hybrid.log; : > $OUT
for c in 2 3 4 6; do PGRT_AUTO_HYBRID_MIN=1 PGRT_TRACE_CTAS_PER_SM=$c timeout 200 python tools/quick_shard.py --ranks 8 --depth 32 --frames 1500 --tag hybrid_tracectas$c 2>&1 | tail -1 | cut -c1-300 >> $OUT; done
PGRT_AUTO_HYBRID_MIN=1 PGRT_TRACE_CTAS_PER_SM=3 PGRT_TRACE_REFILL=16 timeout 200 python tools/quick_shard.py --ranks 8 --depth 32 --frames 1500 --tag hybrid_tracectas3_refill16 2>&1 | tail -1 | cut -c1-300 >> $OUT
timeout 200 python tools/quick_shard.py --ranks 8 --depth 32 --frames 1500 --tag fused 2>&1 | tail -1 | cut -c1-300 >> $OUT
for f in 1 4; do PGRT_FLUSH_CTAS_PER_SM=$f timeout 200 python tools/scale_probe.py --depth 16 --frames 600 --variants solo 2>&1 | grep -E "N=|rror" | sed "s/^/flushctas$f /" >> $OUT; done
cat $OUT
