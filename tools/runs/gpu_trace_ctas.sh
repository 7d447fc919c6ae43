mkdir -p gpurun_out
for t in 6 5 4; do for w in c2 c4; do
  PGRT_TRACE_CTAS_PER_SM=$t timeout 300 python bench.py --workload $w --steps 100 --no-cpu-baseline 2>/dev/null | tail -1 | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print('trace_ctas=$t $w', round(d['value']), 'Mrays/s', round(d['ms_per_step'],4), 'ms e2e', round(d['e2e']['value']))"
done; done
