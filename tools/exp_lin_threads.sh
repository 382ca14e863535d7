#!/bin/bash
# experiment (run under gpurun): CTA size of linearise_fast_kernel -- registers per thread against warps per SM
for T in 384 320 256 224 192; do
  rm -f f16_mpc_oop_py_b200/csrc/build/linearise_fast.o
  make -s -C f16_mpc_oop_py_b200/csrc EXTRA="-DF16_LF_THREADS_FORWARD=$T" > /dev/null 2>&1
  grep -A2 "linearise_fast_kernelILi1" f16_mpc_oop_py_b200/csrc/build/ptxas_linearise_fast.log | grep -E "spill|Used" | tr '\n' ' '; echo
  python bench.py --steps 1 --warmup 0 --euler-steps 10 --no-cpu-baseline --no-cfg5 --e2e-steps 0 | python -c "
import json,sys; b=json.loads(sys.stdin.read()); j=b['jacobians']; print('threads $T central %.4e (%.3f)  forward %.4e (%.3f)' % (j['central']['value'], j['central']['fp64_frac'], j['forward']['value'], j['forward']['fp64_frac']))"
done
