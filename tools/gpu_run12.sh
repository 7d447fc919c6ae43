mkdir -p gpurun_out
python tools/prof_frame.py --workload c2 --frames 3 > gpurun_out/plain5.log 2>&1 && ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_r1_bvh8f.csv python tools/prof_frame.py --workload c2 --frames 3 > gpurun_out/ncu5.log 2>&1
python tools/prof_frame.py --workload c2 --frames 3 > gpurun_out/plain6.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:"k_trace|k_shade|k_resolve" -s 4 -c 3 -f -o gpurun_out/prof_bvh8f_r1 python tools/prof_frame.py --workload c2 --frames 3 > gpurun_out/ncu6.log 2>&1
tail -2 gpurun_out/ncu6.log
