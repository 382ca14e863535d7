#!/bin/bash
# Weak-scaling evidence (run under gpurun --gpus 8): the default bench line (cfg 2 + the cfg-5 leg at 8 Mi aircraft per GPU) at
# 8 and 4 GPUs under torchrun, then ONE process driving all GPUs through the C ABI alone (f16_init_devices): the library bench and
# the plain-C cfg-5 example.  Usage: tools/scale_run.sh <tag>
TAG=${1:-r2x}
O=gpurun_out
for n in 8 4; do
  python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port $((29500 + n)) \
      bench.py --gpus $n --steps 3 --warmup 3 > $O/scale_${TAG}_n$n.json 2>> $O/scale_${TAG}.err
done
python tools/bench_lib_multi.py > $O/lib_multi_${TAG}_n8.jsonl 2>> $O/scale_${TAG}.err
gcc -O2 -Iinclude examples/cfg5_closed_loop_multi_gpu.c -o /tmp/cfg5 -Lf16_mpc_oop_py_b200 -lf16_b200 -Wl,-rpath,$PWD/f16_mpc_oop_py_b200 \
  && /tmp/cfg5 8388608 10000 > $O/c_example_cfg5_${TAG}_n8.log 2>&1
for f in $O/scale_${TAG}_n8.json $O/scale_${TAG}_n4.json; do python - "$f" <<'PY'
import json,sys
for l in open(sys.argv[1]):
    if l.startswith("{"):
        d=json.loads(l); c=d.get("cfg5_lqr") or {}
        print(sys.argv[1], d["n_gpus"], "%.4g"%d["value"], "frac %.3f"%d["roofline"]["frac"], "e2e %.4g"%d["e2e"]["value"], "pageable %.4g"%d["e2e"]["pageable"]["value"],
              "cfg5 %.4g (%d aircraft, alive %.3f)"%(c.get("value",0), c.get("aircraft_total",0), c.get("alive_fraction",0)))
PY
done
cat $O/lib_multi_${TAG}_n8.jsonl | cut -c1-230; cat $O/c_example_cfg5_${TAG}_n8.log; tail -3 $O/scale_${TAG}.err
