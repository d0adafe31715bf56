#!/bin/bash
# One GPU-box job (1 GPU): the evidence of a round.  Everything lands in gpurun_out/ (scratch);
# tools/update_profiles.sh copies the summaries into profiles/.   usage: tools/gpu_checkpoint.sh <round tag, e.g. r2>
R=${1:-r2}
O=gpurun_out
if [ -n "$RUN_TESTS" ]; then (time python -m pytest tests -q -m gpu) > $O/${R}_gpu_tests.txt 2>&1; tail -4 $O/${R}_gpu_tests.txt; fi
timeout 900 python bench.py > $O/${R}_bench_default.json 2> $O/${R}_bench_default.err; tail -c 300 $O/${R}_bench_default.err
timeout 600 python bench.py --steps 20 --warmup 5 > $O/${R}_bench_steps20.json 2> $O/${R}_bench_steps20.err
timeout 600 python bench.py --impl reference --steps 20 --warmup 5 > $O/${R}_bench_reference_arm.json 2> $O/${R}_bench_reference_arm.err
timeout 600 python bench.py --genomes-per-merge 10 --steps 9 --warmup 2 --no-cpu-baseline --no-rank-bench > $O/${R}_bench_g10.json 2> $O/${R}_bench_g10.err
# launch list (shares, not absolutes: cold cache, serialised)
timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"k_|Kernel" -c 6000 --csv --log-file $O/${R}_launches.csv python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-e2e --no-rank-bench > $O/ncu_launches.log 2>&1
python tools/launch_shares.py $O/${R}_launches.csv > $O/${R}_launch_shares.txt
# ncu --set full of the hot kernels
for k in k_walk_pair k_fix_chain k_emit_bm_fast k_write_walk; do
  timeout 400 ncu --set full --clock-control none --import-source on -k regex:"$k" -s 30 -c 1 -f -o $O/prof_$k python bench.py --steps 40 --warmup 3 --no-cpu-baseline --no-e2e --no-rank-bench --no-build > $O/ncu_$k.log 2>&1
  python tools/ncu_summary.py $O/prof_$k.ncu-rep > $O/${R}_ncu_$k.txt 2>&1
  ncu -i $O/prof_$k.ncu-rep --page raw --csv > $O/${R}_ncu_${k}_raw.csv 2>/dev/null
done
timeout 300 ncu --set full --clock-control none --import-source on -k regex:"k_lf_bm" -s 2 -c 1 -f -o $O/prof_k_lf_bm python tools/rank_bench.py --kind bitmap --reps 1 > $O/ncu_lf_bm.log 2>&1
python tools/ncu_summary.py $O/prof_k_lf_bm.ncu-rep > $O/${R}_ncu_k_lf_bm.txt 2>&1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:"k_lf_t1" -s 2 -c 1 -f -o $O/prof_k_lf_t1 python tools/rank_bench.py --kind rle --variants 11 --reps 1 --blocks 2.1e6 --max-len 2000 > $O/ncu_lf_t1.log 2>&1
python tools/ncu_summary.py $O/prof_k_lf_t1.ncu-rep > $O/${R}_ncu_k_lf_t1.txt 2>&1
# rank micro-benchmarks
timeout 300 python tools/rank_bench.py --kind bitmap > $O/${R}_rank_bench_bitmap.jsonl 2>/dev/null
timeout 300 python tools/rank_bench.py --kind rle --variants 8,4,2,22,11,1 > $O/${R}_rank_bench_rle.jsonl 2>/dev/null
timeout 300 python tools/rank_bench.py --kind rle --variants 2,11 --blocks 2.1e6 --max-len 2000 > $O/${R}_rank_bench_rle_1e11_symbols.jsonl 2>/dev/null
# suffix sorter, CLI end to end, C2-style batches
timeout 300 python tools/bwt_bench.py 1 2 4 10 > $O/${R}_bwt_bench.jsonl 2>/dev/null
timeout 600 python tools/cli_e2e.py 24 5000000 > $O/${R}_cli_e2e.json 2> $O/cli_e2e.err
timeout 900 python bench.py --config c2s --c2s-genomes 600 --no-cpu-baseline --no-rank-bench --no-e2e > $O/${R}_bench_c2s.json 2> $O/${R}_bench_c2s.err; tail -c 300 $O/${R}_bench_c2s.err
ls $O | wc -l
