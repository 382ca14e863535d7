#!/bin/bash
# the two headline step measurements, short form (run under gpurun): cfg 2 (2^20 x 10^4 open loop) and cfg 5 (8 Mi closed loop)
python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-jacobians --no-cfg5 --e2e-steps 0 | python -c "
import json,sys; b=json.loads(sys.stdin.read()); print('cfg2 %.4e frac %.4f kernel_ms %.2f' % (b['value'], b['roofline']['frac'], b['roofline']['kernel_ms']))"
python bench.py --steps 2 --warmup 2 --no-cpu-baseline --no-jacobians --no-cfg5 --e2e-steps 0 --workload lqr --aircraft 8388608 | python -c "
import json,sys; b=json.loads(sys.stdin.read()); print('cfg5 %.4e frac %.4f' % (b['value'], b['roofline']['frac']))"
