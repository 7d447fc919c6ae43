set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -8
for d in 1 2 3 4; do timeout 300 python bench.py --steps 30 --warmup 3 --no-cpu-baseline --inflight $d > gpurun_out/bench_inflight$d.log 2>&1; tail -1 gpurun_out/bench_inflight$d.log | cut -c1-330; done
