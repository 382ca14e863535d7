#!/bin/bash
# ncu evidence for the fused step kernel (run under gpurun, one GPU): launch list of the default bench command and one
# --set full capture of a short launch.  Usage: tools/prof_step.sh <tag> [threads]
TAG=${1:-r1}
THR=${2:-256}
set -x
python bench.py --steps 2 --warmup 1 --step-threads $THR > gpurun_out/bench_${TAG}.json 2> gpurun_out/bench_${TAG}.err || exit 1
cat gpurun_out/bench_${TAG}.json
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_${TAG}.csv \
    python bench.py --steps 2 --warmup 1 --step-threads $THR --no-cpu-baseline --no-cfg5 > gpurun_out/ncu_launches_${TAG}.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:step_ -c 1 -f -o gpurun_out/prof_step_${TAG} \
    python bench.py --steps 1 --warmup 0 --euler-steps 200 --step-threads $THR --no-cpu-baseline --no-cfg5 --e2e-steps 0 > gpurun_out/ncu_full_${TAG}.log 2>&1
ls -la gpurun_out/prof_step_${TAG}.ncu-rep
