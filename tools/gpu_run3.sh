set -x
mkdir -p gpurun_out
python tools/prof_frame.py --workload c2 --frames 3 > gpurun_out/plain3.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:"k_trace|k_secondary" -s 4 -c 2 -f -o gpurun_out/prof_bvh8_r1 python tools/prof_frame.py --workload c2 --frames 3 > gpurun_out/ncu4.log 2>&1
tail -3 gpurun_out/ncu4.log
