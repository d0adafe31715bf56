#!/bin/bash
# usage: tools/async_probe.sh <n_gpus> [extra bench args]   -- the multi-device merge with and without the asynchronous exchange + merge
N=$1; shift
for A in 1 0; do
  timeout 240 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $((29520 + A)) bench.py --gpus $N --steps 20 --warmup 5 --no-cpu-baseline --no-rank-bench --no-e2e --param dist_async=$A "$@" > gpurun_out/r2_dist_async${A}_n$N.json 2> gpurun_out/r2_dist_async${A}_n$N.err
  python - <<PY
import json
try:
    d=json.load(open("gpurun_out/r2_dist_async${A}_n$N.json"))
    print("N=$N dist_async=$A value %.4g ms/step %.4g" % (d["value"], d["ms_per_step"]), {k: round(v, 3) for k, v in d["phase_ms_per_step"].items()})
except Exception as e:
    print("N=$N dist_async=$A failed", e)
PY
  tail -3 gpurun_out/r2_dist_async${A}_n$N.err
done
