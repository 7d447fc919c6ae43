mkdir -p gpurun_out
( time python bench.py ) > gpurun_out/bench_default.log 2>&1; tail -4 gpurun_out/bench_default.log | cut -c1-2500
( time python bench.py --impl reference --steps 5 --warmup 1 ) > gpurun_out/bench_reference.log 2>&1; tail -4 gpurun_out/bench_reference.log | cut -c1-900
python -c "import __graft_entry__ as g; g.smoke()"
nproc
