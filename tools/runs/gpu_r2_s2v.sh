#!/bin/bash
# round 2, session 2, one GPU: the top of the tree staged in shared memory for k_trace (-DPGRT_SMEM_TOP=1: the root; 9: the root and the eight nodes behind it)
mkdir -p gpurun_out
OUT=gpurun_out/r2s2_smemtop.log; : > $OUT
timeout 120 python tools/frame_hash.py >> $OUT 2>&1
for v in smemtop1 smemtop9; do PGRT_LIB=$PWD/build/variants/$v.so timeout 120 python tools/frame_hash.py 2>&1 | tail -1 >> $OUT; done
timeout 200 python tools/scale_probe.py --depth 16 --frames 600 --variants solo 2>&1 | grep -E "N=|rror" | sed 's/^/default /' >> $OUT
for v in smemtop1 smemtop9; do PGRT_LIB=$PWD/build/variants/$v.so timeout 200 python tools/scale_probe.py --depth 16 --frames 600 --variants solo 2>&1 | grep -E "N=|rror" | sed "s/^/$v /" >> $OUT; done
timeout 200 python tools/scale_probe.py --depth 16 --frames 600 --variants solo 2>&1 | grep -E "N=|rror" | sed 's/^/default /' >> $OUT
cat $OUT
