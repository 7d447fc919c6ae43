set -x
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -x -q 2>&1 | tail -15
timeout 300 python tools/prof_frame.py --workload c2 --frames 3 --stats gpurun_out/level_stats_r1_dyn.json > gpurun_out/stats_dyn.log 2>&1; cat gpurun_out/stats_dyn.log
timeout 300 python bench.py --steps 20 --warmup 3 --no-cpu-baseline > gpurun_out/bench_dyn.log 2>&1; cat gpurun_out/bench_dyn.log
