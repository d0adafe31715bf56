#!/bin/bash
# usage: tools/sweep.sh <tag> "<args1>" "<args2>" ...   -- short bench runs, one summary line each
tag=$1; shift
i=0
for args in "$@"; do
  python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-rank-bench --no-e2e --no-build $args > gpurun_out/r2_sweep_${tag}_$i.json 2> gpurun_out/r2_sweep_${tag}_$i.err
  python - <<PY
import json
try:
    d=json.load(open("gpurun_out/r2_sweep_${tag}_$i.json"))
    print("[$args] value %.4g ms/step %.4g" % (d["value"], d["ms_per_step"]), {k: round(v, 3) for k, v in d["phase_ms_per_step"].items()}, d["walk"]["last_step_fix_rows"], d["walk"]["last_step_fix_longest_chain"])
except Exception as e:
    print("[$args] failed", e)
PY
  i=$((i+1))
done
