#!/bin/bash
# round 2, session 2, one GPU: regression of the final tree: smoke(), the GPU suite
mkdir -p gpurun_out
timeout 300 python __graft_entry__.py smoke 2>&1 | tail -2
timeout 1500 python -m pytest tests -m gpu -x -q --timeout 280 2>&1 | tail -4
