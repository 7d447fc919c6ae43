mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q -x 2>&1 | tail -4
for cfg in "32 6" "16 6" "12 6" "8 6" "4 6" "12 4" "12 8" "12 12"; do set -- $cfg
r=$(PGRT_TRACE_REFILL=$1 PGRT_TRACE_CTAS_PER_SM=$2 timeout 300 python bench.py --steps 60 --no-cpu-baseline 2>&1 | tail -1 | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print(round(d['value']), round(d['ms_per_step'],4), 'e2e', round(d['e2e']['value']), 'L0', round(d['roofline']['level0_trace_ms'],3), 'sec', round(d['roofline']['secondary_ms'],3), 'unpip', round(d['roofline']['frame_ms_unpipelined'],3))")
echo "refill=$1 ctas=$2 : $r"
done 2>&1 | tee gpurun_out/sweep_refill.log
