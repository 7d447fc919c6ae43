mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -3
for d in 4 6 8; do
  r=$(timeout 300 python bench.py --steps 60 --warmup 3 --no-cpu-baseline --inflight $d 2>&1 | tail -1 | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print(round(d['value']), round(d['ms_per_step'],3), 'e2e', round(d['e2e']['value']), 'L0', round(d['roofline']['level0_trace_ms'],3), 'sec_ms', round(d['roofline']['secondary_ms'],3), 'unpip', round(d['roofline']['frame_ms_unpipelined'],3), d['gpu_launches'])")
  echo "contrib-words inflight=$d : $r"
done 2>&1 | tee gpurun_out/sweep_words.log
