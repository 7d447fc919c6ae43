mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q 2>&1 | tail -3
for v in default s44 s1616 s88r4 s11; do
lib=""; [ $v != default ] && lib=$PWD/build/variants/libpgrt_$v.so
r=$(PGRT_LIB=$lib timeout 300 python bench.py --steps 60 --no-cpu-baseline 2>&1 | tail -1 | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print(round(d['value']), round(d['ms_per_step'],4), 'e2e', round(d['e2e']['value']), 'L0', round(d['roofline']['level0_trace_ms'],3), 'sec', round(d['roofline']['secondary_ms'],3), 'unpip', round(d['roofline']['frame_ms_unpipelined'],3))")
echo "variant=$v : $r"
done 2>&1 | tee gpurun_out/sweep_secondary_sm.log
