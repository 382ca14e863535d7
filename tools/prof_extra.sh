O=gpurun_out
ncu --set full --clock-control none --import-source on -k regex:step_lofi_fast -c 1 -f -o $O/prof_lofi_r1n python tools/bench_entry_points.py --math fast --fi 0 --sizes 1048576 --only step100 --reps 1 > $O/ncu_lofi_r1n.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:partial_kernel --launch-skip 2 -c 2 -f -o $O/prof_stats_r1n python tools/bench_entry_points.py --math fast --sizes 8388608 --only summary --reps 1 > $O/ncu_stats_r1n.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:calc_xdot_kernel --launch-skip 1 -c 1 -f -o $O/prof_xdot_r1n python tools/bench_entry_points.py --math strict --sizes 8388608 --only calc_xdot --reps 1 > $O/ncu_xdot_r1n.log 2>&1
ls -la $O/prof_lofi_r1n.ncu-rep $O/prof_stats_r1n.ncu-rep $O/prof_xdot_r1n.ncu-rep
