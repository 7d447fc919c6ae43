#!/bin/bash
# round 2, session 2, two GPUs: the host-side look at the "consumed" word (8 channels), and 32 channels by default
mkdir -p gpurun_out
OUT=gpurun_out/r2s2_probe_n2c.log; : > $OUT
run() { echo "== $1 $2" >> $OUT; env $1 $2 timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29542 tools/scale_probe.py --depth 16 --frames 500 --variants $3 2>&1 | grep -E "^N=|rror" >> $OUT; }
run CUDA_DEVICE_MAX_CONNECTIONS=8 X=1 counter,words
run CUDA_DEVICE_MAX_CONNECTIONS=8 PGRT_DIST_PROBE=nospin counter
run X=1 X=1 counter,words,allreduce,solo
run X=1 PGRT_DIST_PROBE=nospin counter
cat $OUT
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29531 tools/p2p_check.py > gpurun_out/r2s2_p2p_n2.log 2>&1
grep -E "ranks|Error|error|assert" gpurun_out/r2s2_p2p_n2.log | tail -12
