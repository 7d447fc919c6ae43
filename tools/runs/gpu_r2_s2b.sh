#!/bin/bash
# round 2, session 2, two GPUs: where a sharded frame's time goes
mkdir -p gpurun_out
timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29541 tools/scale_probe.py --depth 8,16 --frames 500 > gpurun_out/r2s2_probe_n2.log 2>&1
grep -E "^N=|rror" gpurun_out/r2s2_probe_n2.log | tail -30
