mkdir -p gpurun_out
run() { # workload inflight steps
  timeout 300 python bench.py --workload $1 --steps $3 --inflight $2 --no-cpu-baseline 2>/dev/null | tail -1 | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print('$1 inflight=$2', round(d['value']), 'Mrays/s', round(d['ms_per_step'],4), 'ms e2e', round(d['e2e']['value']), 'unpip', round(d['roofline']['frame_ms_unpipelined'],3), 'host', round(d['config']['host_issue_us_per_step'],1))"
}
for d in 6 8 12 16; do run c1 $d 120; done
for d in 6 8 12; do run c2 $d 60; done
for d in 6 12; do run c4 $d 60; done
