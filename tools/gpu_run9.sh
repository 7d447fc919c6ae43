mkdir -p gpurun_out
for w in c1 c4 c3 c5; do
  steps=20; [ $w = c3 ] && steps=4; [ $w = c5 ] && steps=6
  timeout 900 python bench.py --workload $w --steps $steps --warmup 3 --no-cpu-baseline > gpurun_out/bench_$w.log 2>&1
  tail -1 gpurun_out/bench_$w.log | python -c "
import json,sys
try:
    d=json.loads(sys.stdin.read()); print('$w', round(d['value']), 'Mrays/s', round(d['ms_per_step'],3), 'ms e2e', round(d['e2e']['value']), d['config']['rays_per_frame'], 'L0', d['roofline']['level0_trace_ms'], 'unpip', d['roofline']['frame_ms_unpipelined'], d['config']['bvh'])
except Exception as e: print('$w FAILED', e)
"
  tail -3 gpurun_out/bench_$w.log | cut -c1-300 | grep -v metric
done 2>&1 | tee gpurun_out/workloads.log
