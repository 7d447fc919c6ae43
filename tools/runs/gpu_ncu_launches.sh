# launch list of the benchmark command (B200_PROFILING.md recipe): plain run first, then the same under ncu
mkdir -p gpurun_out
python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/plain_final2.log 2>&1 || { tail -5 gpurun_out/plain_final2.log; exit 1; }
ncu --metrics gpu__time_duration.sum --clock-control none -c 1500 --csv --log-file gpurun_out/launches_r1_final2.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_final2.log 2>&1
echo "ncu rc=$?"; wc -l gpurun_out/launches_r1_final2.csv; tail -2 gpurun_out/ncu_final2.log | cut -c1-300
