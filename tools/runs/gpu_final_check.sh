# end-of-round check: smoke(), the whole GPU suite, then C2 / C4 / C5 (float-plane and quantised nodes)
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -1
timeout 900 python -m pytest tests -m gpu -q 2>&1 | tail -3
for w in c2 c4 c5; do
  steps=100; [ $w = c5 ] && steps=12
  timeout 300 python bench.py --workload $w --steps $steps --no-cpu-baseline 2>/dev/null | tail -1 | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print('$w', round(d['value']), 'Mrays/s', round(d['ms_per_step'],4), 'ms e2e', round(d['e2e']['value']), 'unpip', round(d['roofline']['frame_ms_unpipelined'],3), 'L0', round(d['roofline']['level0_trace_ms'],3))"
done
