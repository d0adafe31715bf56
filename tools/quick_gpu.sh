#!/bin/bash
# quick GPU check: core parity tests + a short bench; usage: tools/quick_gpu.sh <tag> [bench args]
tag=$1; shift
(time python -m pytest tests/test_gpu_path.py tests/test_gpu_scale.py -x -q -m gpu -k "merge or optional or sharded or scale or insert_multi or transition or pointer" 2>&1 | tail -8) 2>&1
python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-rank-bench "$@" > gpurun_out/r2_bench_$tag.json 2> gpurun_out/r2_bench_$tag.err
python - <<PY
import json
d=json.load(open("gpurun_out/r2_bench_$tag.json"))
print("value %.4g e2e %.4g ms/step %.4g" % (d["value"], d["e2e"]["value"] if d.get("e2e") else 0, d["ms_per_step"]))
print({k: round(v, 4) for k, v in d["phase_ms_per_step"].items()}, d.get("walk"))
PY
tail -3 gpurun_out/r2_bench_$tag.err
