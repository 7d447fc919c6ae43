#!/bin/bash
# round 2, session 2, one GPU: ncu --set full of one C2 frame's kernels (hybrid and fused scheduler); keeper CTAs on an eighth of the frame
mkdir -p gpurun_out
timeout 300 python tools/prof_frame.py --workload c2 --frames 3 > gpurun_out/r2s2_prof_c2.log 2>&1 && \
timeout 900 ncu --set full --clock-control none --import-source on -k regex:'k_trace|k_shade|k_phong|k_frame|k_resolve' --launch-skip 5 -c 5 -f -o gpurun_out/r2s2_ncu_hybrid python tools/prof_frame.py --workload c2 --frames 2 > gpurun_out/r2s2_ncu_h.log 2>&1
tail -2 gpurun_out/r2s2_ncu_h.log | cut -c1-300
timeout 900 ncu --set full --clock-control none --import-source on -k regex:'k_frame' --launch-skip 1 -c 1 -f -o gpurun_out/r2s2_ncu_fused python tools/prof_frame.py --workload c2 --frames 2 --params '{"scheduler": 0}' > gpurun_out/r2s2_ncu_f.log 2>&1
tail -2 gpurun_out/r2s2_ncu_f.log | cut -c1-300
OUT=gpurun_out/r2s2_keep8.log; : > $OUT
for k in 1 2 4 8; do PGRT_KEEP_CTAS=$k timeout 200 python tools/quick_shard.py --ranks 8 --depth 32 --frames 1500 --tag keep$k 2>&1 | tail -1 | cut -c1-300 >> $OUT; done
PGRT_AUTO_HYBRID_MIN=1 PGRT_KEEP_CTAS=2 timeout 200 python tools/quick_shard.py --ranks 8 --depth 32 --frames 1500 --tag hybrid_keep2 2>&1 | tail -1 | cut -c1-300 >> $OUT
PGRT_KEEP_CTAS=2 timeout 200 python tools/quick_shard.py --ranks 4 --depth 16 --frames 1500 --tag keep2 2>&1 | tail -1 | cut -c1-300 >> $OUT
PGRT_KEEP_CTAS=2 timeout 200 python tools/quick_shard.py --ranks 1 --depth 16 --frames 600 --tag keep2 2>&1 | tail -1 | cut -c1-300 >> $OUT
cat $OUT
