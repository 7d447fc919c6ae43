#!/bin/bash
# round 2, session 2, two GPUs: which part of the flag protocol costs 0.075 ms per frame
mkdir -p gpurun_out
OUT=gpurun_out/r2s2_probe_n2b.log; : > $OUT
run() { echo "== $1 $2" >> $OUT; env $1 $2 timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29542 tools/scale_probe.py --depth 16 --frames 500 --variants $3 2>&1 | grep -E "^N=|rror" >> $OUT; }
run PGRT_DIST_PROBE= X=1 counter,solo
run PGRT_DIST_PROBE= CUDA_DEVICE_MAX_CONNECTIONS=32 counter,words,solo
run PGRT_DIST_PROBE=nobefore X=1 counter
run PGRT_DIST_PROBE=nobefore,nomark X=1 counter
run PGRT_DIST_PROBE=nobefore,nomark,nosignal X=1 counter
run PGRT_DIST_PROBE=nosignal X=1 counter
cat $OUT
