set -x
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q 2>&1 | tail -3
python tools/prof_frame.py --workload c2 --frames 3 --stats gpurun_out/level_stats_r1_bvh2.json > gpurun_out/stats.log 2>&1; cat gpurun_out/stats.log
python tools/prof_frame.py --workload c2 --frames 3 > gpurun_out/plain2.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:"k_trace|k_phong|k_shade" -s 66 -c 9 -f -o gpurun_out/prof_bvh2_r1 python tools/prof_frame.py --workload c2 --frames 3 > gpurun_out/ncu3.log 2>&1
tail -3 gpurun_out/ncu3.log
