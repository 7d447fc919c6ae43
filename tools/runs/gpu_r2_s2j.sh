#!/bin/bash
# round 2, session 2, two GPUs: wavefront against fused scheduler with the L2 flush, whole frame (one GPU alone) and half frames
mkdir -p gpurun_out
OUT=gpurun_out/r2s2_sched.log; : > $OUT
CUDA_VISIBLE_DEVICES=0 timeout 200 python tools/scale_probe.py --depth 8,16 --frames 400 --variants solo --params '{"scheduler": 1}' 2>&1 | grep -E "N=|rror" >> $OUT
CUDA_VISIBLE_DEVICES=0 timeout 200 python tools/scale_probe.py --depth 8 --frames 400 --variants solo 2>&1 | grep -E "N=|rror" >> $OUT
timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29542 tools/scale_probe.py --depth 16 --frames 500 --variants counter,solo --params '{"scheduler": 1}' 2>&1 | grep -E "N=|rror" >> $OUT
timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29543 tools/scale_probe.py --depth 16 --frames 500 --variants counter 2>&1 | grep -E "N=|rror" >> $OUT
cat $OUT
