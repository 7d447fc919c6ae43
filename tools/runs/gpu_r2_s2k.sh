#!/bin/bash
# round 2, session 2, one GPU: the hybrid scheduler (level 0 as wavefront kernels, levels >= 1 through the pool): GPU suite, whole frame and shards
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q --timeout 280 2>&1 | tail -6
OUT=gpurun_out/r2s2_hybrid.log; : > $OUT
timeout 200 python tools/scale_probe.py --depth 8,16 --frames 400 --variants solo --params '{"scheduler": 2}' 2>&1 | grep -E "N=|rror" >> $OUT
for r in 2 4 8; do
  d=16; [ $r = 8 ] && d=32
  for s in 0 1 2; do timeout 200 python tools/quick_shard.py --ranks $r --depth $d --frames 1200 --params "{\"scheduler\": $s}" 2>&1 | tail -1 | cut -c1-300 >> $OUT; done
done
cat $OUT
