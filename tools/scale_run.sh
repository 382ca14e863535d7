#!/bin/bash
# Weak-scaling evidence (run under gpurun --gpus 8): the bench line at 1, 2, 4, 8 GPUs (cfg 2) and cfg 5 (closed loop, 8 Mi aircraft per GPU) at 8.
TAG=${1:-r1x}
O=gpurun_out
python bench.py --gpus 1 --no-cpu-baseline > $O/scale_${TAG}_n1.json 2> $O/scale_${TAG}.err
for n in 2 4 8; do
  python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port $((29500 + n)) \
      bench.py --gpus $n --steps 3 --warmup 3 > $O/scale_${TAG}_n$n.json 2>> $O/scale_${TAG}.err
done
python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29520 \
    bench.py --gpus 8 --steps 1 --warmup 3 --workload lqr --aircraft 8388608 --no-jacobians > $O/scale_${TAG}_cfg5_n8.json 2>> $O/scale_${TAG}.err
for f in $O/scale_${TAG}_n*.json $O/scale_${TAG}_cfg5_n8.json; do python - "$f" <<'PY'
import json,sys
for l in open(sys.argv[1]):
    if l.startswith("{"):
        d=json.loads(l); print(sys.argv[1], d["n_gpus"], "%.4g"%d["value"], "frac %.3f"%d["roofline"]["frac"], "e2e %.4g"%d["e2e"]["value"], d["config"]["alive_fraction"])
PY
done
