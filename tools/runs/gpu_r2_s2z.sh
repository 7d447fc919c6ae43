#!/bin/bash
# round 2, session 2, one GPU: PLOC radius 8 (was 16): the GPU suite (the GPU builder must reproduce the emulator's tree), then build and frame times
mkdir -p gpurun_out
timeout 200 python -m pytest tests -m gpu -x -q --timeout 120 2>&1 | tail -3 | tee gpurun_out/r2s2_plocr8_pytest.log
timeout 60 python tools/scale_probe.py --depth 16 --frames 400 --variants solo 2>&1 | grep -E "N=|rror" | tee gpurun_out/r2s2_plocr8.log
timeout 60 python tools/build_only.py --workload c5 --reps 2 2>&1 | tail -1 | cut -c1-300 | tee -a gpurun_out/r2s2_plocr8.log
