# full GPU regression: parity suite, then the five workloads (value, ms/step, build ms)
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q 2>&1 | tail -6
for w in c2 c1 c4 c5 c3; do
  steps=60; [ $w = c3 ] && steps=4; [ $w = c5 ] && steps=12
  timeout 400 python bench.py --workload $w --steps $steps --no-cpu-baseline 2>/dev/null | tail -1 | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print('$w', round(d['value']), 'Mrays/s', round(d['ms_per_step'],4), 'ms  e2e', round(d['e2e']['value']), ' build', {k: (round(v,2) if isinstance(v,float) else v) for k,v in d['config']['bvh'].items() if k in ('nodes','build_ms','sah_cost','depth','ploc_passes','node_bytes')}, d['clocks'])"
done
