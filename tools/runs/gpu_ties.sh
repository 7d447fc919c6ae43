mkdir -p gpurun_out
for t in 3 0; do for w in c4 c2 c5; do
  steps=60; [ $w = c5 ] && steps=12
  PGRT_PLOC_TIES=$t timeout 400 python bench.py --workload $w --steps $steps --no-cpu-baseline 2>/dev/null | tail -1 | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print('ties=$t $w', round(d['value']), 'Mrays/s', round(d['ms_per_step'],4), {k: (round(v,2) if isinstance(v,float) else v) for k,v in d['config']['bvh'].items() if k in ('nodes','sah_cost','depth','ploc_passes')}, )"
done; done
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -q -k "awkward or non_finite or soup or structure" 2>&1 | tail -4
