#!/bin/bash
# round 2, session 2, one GPU: k_trace grid sized to the rays of level 0 (one CTA per 1024 / 2048 rays), whole frame and shards
mkdir -p gpurun_out
OUT=gpurun_out/r2s2_tracegrid.log; : > $OUT
for rp in 1024 2048; do
PGRT_TRACE_RAYS_PER_CTA=$rp timeout 200 python tools/scale_probe.py --depth 16 --frames 600 --variants solo 2>&1 | grep -E "N=|rror" | sed "s/^/raysPerCta$rp /" >> $OUT
for r in 2 4; do PGRT_TRACE_RAYS_PER_CTA=$rp timeout 200 python tools/quick_shard.py --ranks $r --depth 16 --frames 1200 --tag raysPerCta$rp 2>&1 | tail -1 | cut -c1-300 >> $OUT; done
PGRT_TRACE_RAYS_PER_CTA=$rp PGRT_AUTO_HYBRID_MIN=1 timeout 200 python tools/quick_shard.py --ranks 8 --depth 32 --frames 1500 --tag hybrid_raysPerCta$rp 2>&1 | tail -1 | cut -c1-300 >> $OUT
done
cat $OUT
