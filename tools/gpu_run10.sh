mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -5
for f in 1 0; do
  r=$(PGRT_FUSE_RAYGEN=$f timeout 300 python bench.py --steps 40 --warmup 3 --no-cpu-baseline 2>&1 | tail -1 | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print(round(d['value']), round(d['ms_per_step'],3), 'e2e', round(d['e2e']['value']), 'L0', round(d['roofline']['level0_trace_ms'],3), 'sec_ms', round(d['roofline']['secondary_ms'],3), 'unpip', round(d['roofline']['frame_ms_unpipelined'],3), d['gpu_launches'])")
  echo "fuse_raygen=$f : $r"
done 2>&1 | tee gpurun_out/sweep_fuse.log
