#!/bin/bash
# round 2: the claim-patience rule of k_frame's tail; A/B of builds and knobs; ncu of the primary phase alone and of the whole frame
mkdir -p gpurun_out
OUT=gpurun_out/r2_tail.log; : > $OUT
for pat in 1 4 16; do PGRT_CLAIM_PATIENCE=$pat timeout 120 python tools/quick_c2.py --tag patience$pat >> $OUT 2>&1; done
for lib in build/variants/*.so; do PGRT_LIB=$PWD/$lib timeout 120 python tools/quick_c2.py --tag $(basename $lib .so) >> $OUT 2>&1; done
for keep in 2 5; do PGRT_KEEP_CTAS_PER_SM=$keep timeout 120 python tools/quick_c2.py --tag keep$keep >> $OUT 2>&1; done
for d in 1 2 8; do timeout 120 python tools/quick_c2.py --depth $d --tag depth$d >> $OUT 2>&1; done
cat $OUT
timeout 600 python -m pytest tests -m gpu -x -q --timeout 240 2>&1 | tail -25 > gpurun_out/r2_pytest.log
tail -4 gpurun_out/r2_pytest.log
timeout 300 python tools/prof_frame.py --workload c2 --frames 3 --params '{"shader_mode": 1}' > gpurun_out/r2_prof_lambert.log 2>&1 && \
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_frame -c 1 -f -o gpurun_out/r2_ncu_kframe_lambert python tools/prof_frame.py --workload c2 --frames 2 --params '{"shader_mode": 1}' > gpurun_out/r2_ncu1.log 2>&1
cat gpurun_out/r2_prof_lambert.log | head -3
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_frame -c 1 -f -o gpurun_out/r2_ncu_kframe python tools/prof_frame.py --workload c2 --frames 2 > gpurun_out/r2_ncu.log 2>&1
tail -2 gpurun_out/r2_ncu.log
