#!/bin/bash
# Copies the judged summaries of a round from gpurun_out/ (scratch) into profiles/ (tracked).
# usage: tools/update_profiles.sh <round tag, e.g. r2>      (after tools/gpu_checkpoint.sh <tag> ran on the GPU box)
set -e
cd "$(dirname "$0")/.."
R=${1:-r2}
O=gpurun_out
for f in bench_default bench_steps20 bench_reference_arm bench_g10 bench_c2s cli_e2e \
         scale_n2 scale_n4 scale_n8 strong_n1 strong_n2 strong_n4 strong_n8; do
  [ -s $O/${R}_$f.json ] && cp $O/${R}_$f.json profiles/${R}_$f.json
done
for f in launches.csv launch_shares.txt rank_bench_bitmap.jsonl rank_bench_rle.jsonl rank_bench_rle_1e11_symbols.jsonl bwt_bench.jsonl fmd_dev_bench.json; do
  [ -s $O/${R}_$f ] && cp $O/${R}_$f profiles/${R}_$f
done
for k in k_walk_pair k_fix_chain k_emit_bm_fast k_write_walk k_lf_bm k_lf_t1; do
  [ -s $O/${R}_ncu_$k.txt ] && cp $O/${R}_ncu_$k.txt profiles/${R}_ncu_$k.txt
  [ -s $O/${R}_ncu_${k}_raw.csv ] && cp $O/${R}_ncu_${k}_raw.csv profiles/${R}_ncu_${k}_raw.csv
done
# SASS of the hot kernels from the built library (what the GPU box ran): mnemonic histogram + full listing of each
python tools/sass_excerpts.py ropebwt3_b200/librb3b200.so profiles/${R}_sass
ls profiles | grep "^${R}_" | wc -l
