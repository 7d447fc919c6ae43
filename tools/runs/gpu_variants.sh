# A/B of library variants built under build/variants (PGRT_LIB selects the .so the ctypes binding loads)
mkdir -p gpurun_out
for v in base $(ls build/variants | sed 's/libpgrt_//; s/.so//'); do
  lib=build/variants/libpgrt_$v.so; [ $v = base ] && lib=pgi_raytracing_b200/libpgrt_b200.so
  for w in ${WORKLOADS:-c2}; do
  steps=100; [ $w = c5 ] && steps=12
  PGRT_LIB=$PWD/$lib timeout 300 python bench.py --workload $w --steps $steps --no-cpu-baseline 2>/dev/null | tail -1 | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print('$v $w', round(d['value']), 'Mrays/s', round(d['ms_per_step'],4), 'ms e2e', round(d['e2e']['value']), 'unpip', round(d['roofline']['frame_ms_unpipelined'],3))"
  done
done
