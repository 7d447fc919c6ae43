mkdir -p gpurun_out
for ctas in 3 4; do for d in 2 3 4; do
  r=$(PGRT_SECONDARY_CTAS_PER_SM=$ctas timeout 300 python bench.py --steps 40 --warmup 3 --no-cpu-baseline --inflight $d 2>&1 | tail -1 | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print(round(d['value']), round(d['ms_per_step'],3), 'e2e', round(d['e2e']['value']), 'sec_ms', round(d['roofline']['secondary_ms'],3), 'unpip', round(d['roofline']['frame_ms_unpipelined'],3))")
  echo "ctas=$ctas inflight=$d : $r"
done; done 2>&1 | tee gpurun_out/sweep_sec2.log
