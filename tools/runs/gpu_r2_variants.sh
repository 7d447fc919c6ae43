#!/bin/bash
# round 2: GPU tests (with a per-test timeout), A/B of k_frame builds and knobs, one ncu --set full capture of k_frame
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -x -q --timeout 240 2>&1 | tail -25 > gpurun_out/r2_pytest.log
tail -4 gpurun_out/r2_pytest.log
OUT=gpurun_out/r2_variants.log; : > $OUT
for lib in build/variants/*.so; do
  PGRT_LIB=$PWD/$lib timeout 120 python tools/quick_c2.py --tag $(basename $lib .so) >> $OUT 2>&1
done
for keep in 2 5; do PGRT_KEEP_CTAS_PER_SM=$keep timeout 120 python tools/quick_c2.py >> $OUT 2>&1; done
for claim in 1 8; do PGRT_MIN_CLAIM=$claim timeout 120 python tools/quick_c2.py >> $OUT 2>&1; done
for d in 1 2 8; do timeout 120 python tools/quick_c2.py --depth $d >> $OUT 2>&1; done
cat $OUT
timeout 300 python tools/prof_frame.py --workload c2 --frames 3 > gpurun_out/r2_prof_plain.log 2>&1 && \
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_frame -c 1 -f -o gpurun_out/r2_ncu_kframe python tools/prof_frame.py --workload c2 --frames 2 > gpurun_out/r2_ncu.log 2>&1
tail -3 gpurun_out/r2_ncu.log
timeout 400 python bench.py --steps 20 --warmup 5 --min-seconds 0.5 --no-extra > gpurun_out/r2_bench_new.log 2>&1; tail -c 3000 gpurun_out/r2_bench_new.log
timeout 200 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
