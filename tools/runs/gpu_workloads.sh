# the workloads given as arguments (default c4 c5), twice each, without the parity suite
mkdir -p gpurun_out
for rep in 1 2; do
for w in ${@:-c4 c5}; do
  steps=60; [ $w = c3 ] && steps=4; [ $w = c5 ] && steps=12
  timeout 400 python bench.py --workload $w --steps $steps --no-cpu-baseline 2>/dev/null | tail -1 | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print('$w', round(d['value']), 'Mrays/s', round(d['ms_per_step'],4), 'ms  e2e', round(d['e2e']['value']), {k: (round(v,2) if isinstance(v,float) else v) for k,v in d['config']['bvh'].items() if k in ('nodes','build_ms','sah_cost','depth','ploc_passes','node_bytes')}, d['clocks'])"
done; done
