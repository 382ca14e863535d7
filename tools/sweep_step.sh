#!/bin/bash
# step_kernel configuration sweep on one B200 (run under gpurun): math mode x CTA size x table staging
E=${1:-1000}
MODES=${2:-"strict fast"}
THREADS=${3:-"256 384 512 640 768 1024"}
fmt='import sys,json; d=json.loads(sys.stdin.readline()); r=d["roofline"]; print(sys.argv[1], "%.3e steps/s" % d["value"], "kernel_ms %.1f" % r["kernel_ms"], "fp64 frac %.3f of %.1f TF" % (r["frac"], r["peak"]), "clk", d["clocks"]["sm_mhz"], d["clocks"]["reasons"], "alive", d["config"]["alive_fraction"])'
for math in $MODES; do
  for thr in $THREADS; do
    python bench.py --steps 2 --warmup 1 --euler-steps $E --math $math --step-threads $thr --no-cpu-baseline --no-cfg5 --no-jacobians --e2e-steps 0 | python -c "$fmt" "$math thr=$thr tma"
  done
  python bench.py --steps 2 --warmup 1 --euler-steps $E --math $math --no-table-staging --no-cpu-baseline --no-cfg5 --no-jacobians --e2e-steps 0 | python -c "$fmt" "$math thr=256 l2 "
done
