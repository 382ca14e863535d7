#!/bin/bash
# A/B of step-kernel builds on one GPU box (run under gpurun): for every variants/<name>/libf16_b200.so (built with
# `make -C f16_mpc_oop_py_b200/csrc OUT=$PWD/variants/<name> OBJ=$PWD/variants/<name>/build EXTRA=-D...`) put it in place of the
# package's library and time cfg 2 with the time-chunked kernel and with the plain kernel on a batch without a grid tail.
# The box's copy of the repo is scratch; nothing is written back.   (variants/ is git- and gpurun-ignored: remove the line from
# .gpurunignore while experimenting.)
cp f16_mpc_oop_py_b200/libf16_b200.so /tmp/libf16_b200.so.orig
for d in variants/*/; do
  n=$(basename $d)
  [ -f $d/libf16_b200.so ] || continue
  cp $d/libf16_b200.so f16_mpc_oop_py_b200/libf16_b200.so
  for rep in 1 2; do
  python bench.py --steps 3 --warmup 2 --no-cpu-baseline --no-jacobians --no-cfg5 --e2e-steps 0 | python -c "
import json,sys; b=json.loads(sys.stdin.read()); print('$n chunked 2^20   %.4e frac %.4f kernel_ms %.2f' % (b['value'], b['roofline']['frac'], b['roofline']['kernel_ms']))"
  done
  F16_STEP_CHUNKING=0 python bench.py --steps 3 --warmup 2 --no-cpu-baseline --no-jacobians --no-cfg5 --e2e-steps 0 --aircraft 1022976 | python -c "
import json,sys; b=json.loads(sys.stdin.read()); print('$n plain 18 rounds %.4e frac %.4f kernel_ms %.2f' % (b['value'], b['roofline']['frac'], b['roofline']['kernel_ms']))"
  python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-jacobians --no-cfg5 --e2e-steps 0 --workload lqr --aircraft 2097152 | python -c "
import json,sys; b=json.loads(sys.stdin.read()); print('$n cfg5 2 Mi (plain kernel) %.4e frac %.4f' % (b['value'], b['roofline']['frac']))"
done
cp /tmp/libf16_b200.so.orig f16_mpc_oop_py_b200/libf16_b200.so
