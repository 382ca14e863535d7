#!/bin/bash
# Evidence pass for profiles/ (run under gpurun, one GPU): smoke, GPU test-suite, default bench line (+ reference arm), entry points,
# ncu launch list of the bench command, one --set full capture each of the step kernel at the bench's K, of linearise and of the
# one-shot kernels.  Usage: tools/prof_round.sh <tag>
TAG=${1:-r1x}
O=gpurun_out
python -c "import __graft_entry__ as g; g.smoke()" > $O/smoke_${TAG}.log 2>&1; echo "smoke rc=$?"; tail -2 $O/smoke_${TAG}.log
python -m pytest tests -m gpu -x -q > $O/pytest_${TAG}.log 2>&1; echo "pytest rc=$?"; tail -2 $O/pytest_${TAG}.log
python bench.py > $O/bench_${TAG}.json 2> $O/bench_${TAG}.err || exit 1
python bench.py --impl reference --steps 2 --warmup 1 > $O/bench_ref_${TAG}.json 2>> $O/bench_${TAG}.err
python tools/bench_entry_points.py --math fast > $O/entry_fast_${TAG}.jsonl 2>&1
python tools/bench_entry_points.py --math strict > $O/entry_strict_${TAG}.jsonl 2>&1
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $O/launches_${TAG}.csv \
    python bench.py --no-cpu-baseline --no-cfg5 > $O/ncu_launches_${TAG}.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:step_ -c 1 -f -o $O/prof_step_${TAG} \
    python bench.py --steps 1 --warmup 0 --no-cpu-baseline --no-cfg5 --e2e-steps 0 --no-jacobians > $O/ncu_full_step_${TAG}.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:linearise --launch-skip 3 --launch-count 2 -f -o $O/prof_lin_${TAG} \
    python bench.py --steps 1 --warmup 0 --euler-steps 10 --no-cpu-baseline --no-cfg5 --e2e-steps 0 > $O/ncu_full_lin_${TAG}.log 2>&1
bash tools/prof_oneshot.sh ${TAG} fast > $O/prof_oneshot_${TAG}.log 2>&1
ls -la $O/*${TAG}*
