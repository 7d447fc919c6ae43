mkdir -p gpurun_out
run() { # workload ctas steps
  PGRT_FLUSH_CTAS_PER_SM=$2 timeout 300 python bench.py --workload $1 --steps $3 --no-cpu-baseline 2>/dev/null | tail -1 | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print('$1 flush_ctas=$2', round(d['value']), 'Mrays/s', round(d['ms_per_step'],4), 'ms e2e', round(d['e2e']['value']), 'unpip', round(d['roofline']['frame_ms_unpipelined'],3), 'host', round(d['config']['host_issue_us_per_step'],1))"
}
for c in 8 2 1; do run c1 $c 120; run c2 $c 60; done
run c4 2 60; run c5 2 12
