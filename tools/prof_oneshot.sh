#!/bin/bash
# ncu --set full of the one-shot kernels at 2^20 aircraft (run under gpurun, one GPU).  Usage: tools/prof_oneshot.sh <tag> [math]
TAG=${1:-r2x}; MATH=${2:-fast}
O=gpurun_out
python tools/bench_entry_points.py --math $MATH --sizes 1048576 --only calc_xdot,nlplant,step1,summary > $O/entry_oneshot_${MATH}_${TAG}.jsonl 2>&1
ncu --set full --clock-control none --import-source on -k regex:'calc_xdot|xdot_fast' --launch-skip 1 -c 1 -f -o $O/prof_xdot_${MATH}_${TAG} \
    python tools/bench_entry_points.py --math $MATH --sizes 1048576 --only calc_xdot --reps 1 > $O/ncu_xdot_${MATH}_${TAG}.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:'nlplant|xdot_fast' --launch-skip 1 -c 1 -f -o $O/prof_nlplant_${MATH}_${TAG} \
    python tools/bench_entry_points.py --math $MATH --sizes 1048576 --only nlplant --reps 1 > $O/ncu_nlplant_${MATH}_${TAG}.log 2>&1
ls -la $O/*${TAG}*
