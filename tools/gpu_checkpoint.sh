#!/bin/bash
# One GPU-box job: the full bench (both arms), the ncu launch list and the ncu --set full captures.
# Everything lands in gpurun_out/ (scratch); tools/update_profiles.sh copies the summaries into profiles/.
# RUN_TESTS=1 also runs the parity tests first.
if [ -n "$RUN_TESTS" ]; then python -m pytest tests -x -q -m gpu 2>&1 | tail -3; fi
timeout 900 python bench.py > gpurun_out/bench_full.json 2> gpurun_out/bench_full.err; tail -c 400 gpurun_out/bench_full.err
python - <<'PY'
import json
d=json.load(open("gpurun_out/bench_full.json"))
for k in ["value","ms_per_step","e2e","gpu_launches","clocks","phase_ms_per_step","wall_ms_per_step","setup","index","walk","cpu_baseline","rank_kernel"]: print(k, d.get(k))
print(d["roofline"])
PY
timeout 300 python bench.py --genomes-per-merge 10 --steps 9 --warmup 3 --no-cpu-baseline --no-rank-bench > gpurun_out/bench_g10.json 2> gpurun_out/bench_g10.err; cut -c1-200 gpurun_out/bench_g10.json
timeout 600 python bench.py --impl reference --steps 10 --warmup 3 --ref-budget-s 40 > gpurun_out/bench_ref.json 2> gpurun_out/bench_ref.err; cut -c1-300 gpurun_out/bench_ref.json
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"k_|Kernel" -c 4000 --csv --log-file gpurun_out/launches.csv python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-e2e --no-rank-bench > gpurun_out/ncu_b.log 2>&1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:"k_walk_first" -s 60 -c 1 -o gpurun_out/prof_walk_bm python bench.py --steps 64 --warmup 3 --no-cpu-baseline --no-e2e --no-rank-bench > gpurun_out/ncu_walk.log 2>&1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:"k_walk_fix_log" -s 60 -c 1 -o gpurun_out/prof_fix_bm python bench.py --steps 64 --warmup 3 --no-cpu-baseline --no-e2e --no-rank-bench > gpurun_out/ncu_fix.log 2>&1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:"k_lf_bm" -s 2 -c 1 -o gpurun_out/prof_lf_bm python tools/rank_bench.py --kind bitmap --reps 1 > gpurun_out/ncu_lf.log 2>&1
timeout 300 python tools/rank_bench.py --kind bitmap > gpurun_out/rank_bench_bm.jsonl 2>/dev/null
timeout 300 python tools/rank_bench.py --kind rle > gpurun_out/rank_bench_rle.jsonl 2>/dev/null
ls gpurun_out | wc -l
