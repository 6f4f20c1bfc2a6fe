#!/bin/bash
# Scaling sweep on one 8-GPU box: bench.py at N = 1, 2, 4, 8 back to back (the way the driver does), lines under gpurun_out/.
# usage: tools/scaling_sweep.sh <tag> [extra bench.py flags]
tag=$1; shift
port=29600
for n in 1 2 4 8; do
  port=$((port + 1))
  if [ $n = 1 ]; then
    timeout 600 python bench.py --gpus 1 --steps 30 --warmup 3 "$@" > gpurun_out/${tag}_scale_n1.json 2> gpurun_out/${tag}_scale_n1.err
  else
    timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port $port bench.py --gpus $n --steps 30 --warmup 3 "$@" > gpurun_out/${tag}_scale_n$n.json 2> gpurun_out/${tag}_scale_n$n.err
  fi
  echo "N=$n exit $?"
done
python - <<PY
import json
base = None
for n in (1, 2, 4, 8):
    try:
        d = json.load(open("gpurun_out/${tag}_scale_n%d.json" % n))
    except Exception as e:
        print(n, "no line:", e); continue
    base = base or d["value"]
    print("N=%d value %.4g (%.3f of N x N=1) ms/step %.3f e2e %.4g clocks %s" % (n, d["value"], d["value"] / (n * base), d["ms_per_step"], d["e2e"]["value"], d["clocks"]))
PY
