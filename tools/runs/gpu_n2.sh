mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -q -k "rank or host_memory" 2>&1 | tail -4
run() { n=$1; w=$2; steps=$3; d=$4
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 2952$n bench.py --gpus $n --workload $w --steps $steps --warmup 3 --inflight $d > gpurun_out/bench_${w}_n${n}_d$d.log 2>&1
tail -1 gpurun_out/bench_${w}_n${n}_d$d.log | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print('$w N=$n depth=$d', round(d['value']), 'Mrays/s', round(d['ms_per_step'],4), 'ms e2e', round(d['e2e']['value']), d['config']['gather'], 'host_issue_us', round(d['config']['host_issue_us_per_step'],1), {k: round(v,1) for k,v in d['config']['host_issue_parts_us'].items()}, 'unpip', round(d['roofline']['frame_ms_unpipelined'],3))" || tail -5 gpurun_out/bench_${w}_n${n}_d$d.log
}
for d in ${DEPTHS:-8 16}; do run ${NG:-2} c2 100 $d; done
