mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q 2>&1 | tail -5
python tools/prof_frame.py --workload c2 --frames 3 > gpurun_out/plain7.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:"k_trace|k_secondary" -s 4 -c 2 -f -o gpurun_out/prof_bvh8w_r1 python tools/prof_frame.py --workload c2 --frames 3 > gpurun_out/ncu7.log 2>&1
tail -2 gpurun_out/ncu7.log
