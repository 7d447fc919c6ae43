mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q 2>&1 | tail -4
for g in 1 0; do
r=$(PGRT_GRAPHS=$g timeout 300 python bench.py --steps 60 --no-cpu-baseline 2>&1 | tail -1 | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print(round(d['value']), round(d['ms_per_step'],4), 'e2e', round(d['e2e']['value']), round(d['config']['host_issue_us_per_step'],1), {k: round(v,1) for k,v in d['config']['host_issue_parts_us'].items()})")
echo "graphs=$g: $r"
done
