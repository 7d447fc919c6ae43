#!/bin/bash
# round 2, session 2, one GPU: the three schedulers on the other named workloads
mkdir -p gpurun_out
OUT=gpurun_out/r2s2_sched_workloads.log; : > $OUT
for w in c1 c4; do for s in 0 1 2; do timeout 300 python tools/quick_c2.py --workload $w --frames 200 --depth 8 --tag ${w}_sched$s --params "{\"scheduler\": $s}" 2>&1 | tail -1 | cut -c1-330 >> $OUT; done; done
for s in 0 1 2; do timeout 300 python tools/quick_c2.py --workload c3 --frames 6 --depth 2 --tag c3_sched$s --params "{\"scheduler\": $s}" 2>&1 | tail -1 | cut -c1-330 >> $OUT; done
for s in 0 1 2; do timeout 300 python tools/quick_c2.py --workload c5 --frames 12 --depth 2 --tag c5_sched$s --params "{\"scheduler\": $s}" 2>&1 | tail -1 | cut -c1-330 >> $OUT; done
cat $OUT
