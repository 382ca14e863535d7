#!/bin/bash
# ncu --set full of linearise_batch at 2^17 points (run under gpurun).  Usage: tools/prof_lin.sh <tag> [math]
TAG=${1:-r2x}; MATH=${2:-fast}
O=gpurun_out
ncu --set full --clock-control none --import-source on -k regex:linearise --launch-skip 2 --launch-count 2 -f -o $O/prof_lin_${MATH}_${TAG} \
    python bench.py --steps 1 --warmup 0 --euler-steps 10 --no-cpu-baseline --no-cfg5 --e2e-steps 0 --math $MATH > $O/ncu_lin_${MATH}_${TAG}.log 2>&1
ls -la $O/prof_lin_${MATH}_${TAG}.ncu-rep
