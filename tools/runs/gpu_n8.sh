# C2 under torch.distributed.run at 8, 4 and 2 GPUs of one box, bench defaults (frames in flight chosen by bench.py)
mkdir -p gpurun_out
run() { n=$1; w=$2; steps=$3
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 2952$n bench.py --gpus $n --workload $w --steps $steps --warmup 3 > gpurun_out/bench_${w}_n${n}.log 2>&1
tail -1 gpurun_out/bench_${w}_n${n}.log | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print('$w N=$n depth', d['config']['frames_in_flight'], round(d['value']), 'Mrays/s', round(d['ms_per_step'],4), 'ms e2e', round(d['e2e']['value']), d['e2e'].get('path','')[:24], d['config']['gather'], 'host_issue_us', round(d['config']['host_issue_us_per_step'],1), {k: round(v,1) for k,v in d['config']['host_issue_parts_us'].items()}, 'unpip', round(d['roofline']['frame_ms_unpipelined'],3), d['clocks'])" || tail -5 gpurun_out/bench_${w}_n${n}.log
}
run 8 c2 200
run 4 c2 200
run 2 c2 200
