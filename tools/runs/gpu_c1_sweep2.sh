mkdir -p gpurun_out
run() { # trace_ctas sec_ctas inflight
  PGRT_TRACE_CTAS_PER_SM=$1 PGRT_SECONDARY_CTAS_PER_SM=$2 timeout 300 python bench.py --workload c1 --steps 160 --inflight $3 --no-cpu-baseline 2>/dev/null | tail -1 | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print('c1 trace=$1 sec=$2 inflight=$3', round(d['value']), 'Mrays/s', round(d['ms_per_step'],4), 'ms unpip', round(d['roofline']['frame_ms_unpipelined'],3), 'L0', round(d['roofline']['level0_trace_ms'],3), 'sec', round(d['roofline']['secondary_ms'],3), 'host', round(d['config']['host_issue_us_per_step'],1))"
}
run 1 1 16; run 2 1 16; run 3 1 16; run 6 1 16; run 1 2 16; run 2 2 16
