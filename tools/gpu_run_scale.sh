mkdir -p gpurun_out
for n in 8 4; do
python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 2952$n bench.py --gpus $n --steps 60 --warmup 3 > gpurun_out/bench_n$n.log 2>&1
tail -1 gpurun_out/bench_n$n.log | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print('N=$n', round(d['value']), round(d['ms_per_step'],4), 'e2e', round(d['e2e']['value']), d['config']['gather'], 'unpip', round(d['roofline']['frame_ms_unpipelined'],3), 'sec', round(d['roofline']['secondary_ms'],3), 'L0', round(d['roofline']['level0_trace_ms'],3))" || tail -5 gpurun_out/bench_n$n.log
done
