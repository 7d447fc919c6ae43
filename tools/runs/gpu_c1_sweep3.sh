mkdir -p gpurun_out
run() { # workload inflight steps
  timeout 300 python bench.py --workload $1 --steps $3 --inflight $2 --no-cpu-baseline 2>/dev/null | tail -1 | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print('$1 inflight=$2', round(d['value']), 'Mrays/s', round(d['ms_per_step'],4), 'ms e2e', round(d['e2e']['value']), 'unpip', round(d['roofline']['frame_ms_unpipelined'],3), 'L0', round(d['roofline']['level0_trace_ms'],3), 'sec', round(d['roofline']['secondary_ms'],3))"
}
run c1 16 160; run c2 12 60; run c2 16 60; run c2 8 60; run c3 0 4; run c4 0 60; run c5 0 12
