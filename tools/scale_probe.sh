#!/bin/bash
# usage: tools/scale_probe.sh N [bench args]  -- one torchrun bench at N GPUs, summary line
N=$1; shift
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus $N --steps 20 --warmup 5 --no-cpu-baseline --no-rank-bench "$@" > gpurun_out/r2_scale_n$N.json 2> gpurun_out/r2_scale_n$N.err
python - <<PY
import json
try:
    d=json.load(open("gpurun_out/r2_scale_n$N.json"))
    print("N=$N value %.4g e2e %.4g ms/step %.4g" % (d["value"], d["e2e"]["value"] if d.get("e2e") else 0, d["ms_per_step"]), {k: round(v, 3) for k, v in d["phase_ms_per_step"].items()})
except Exception as e:
    print("N=$N failed", e); print(open("gpurun_out/r2_scale_n$N.err").read()[-1500:])
PY
