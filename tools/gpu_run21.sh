mkdir -p gpurun_out
python bench.py --steps 2 --warmup 3 --no-cpu-baseline --inflight 1 > gpurun_out/plain_bench.log 2>&1 && ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/launches_r1_final.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline --inflight 1 > gpurun_out/ncu_bench.log 2>&1
tail -1 gpurun_out/ncu_bench.log | cut -c1-200
python tools/prof_frame.py --workload c2 --frames 3 > gpurun_out/plain8.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:"k_trace|k_secondary|k_shade|k_phong" -s 8 -c 4 -f -o gpurun_out/prof_final_r1 python tools/prof_frame.py --workload c2 --frames 3 > gpurun_out/ncu8.log 2>&1
tail -2 gpurun_out/ncu8.log
