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
for f in gpu_tests.txt launches.csv launch_shares.txt rank_bench_bitmap.jsonl rank_bench_rle.jsonl rank_bench_rle_1e11_symbols.jsonl bwt_bench.jsonl fmd_dev_bench.json; do
  [ -s $O/${R}_$f ] && cp $O/${R}_$f profiles/${R}_$f
done
for k in k_walk_pair k_fix_chain k_emit_bm_fast k_write_walk k_lf_bm k_lf_t1; do
  [ -s $O/${R}_ncu_$k.txt ] && cp $O/${R}_ncu_$k.txt profiles/${R}_ncu_$k.txt
  [ -s $O/${R}_ncu_${k}_raw.csv ] && cp $O/${R}_ncu_${k}_raw.csv profiles/${R}_ncu_${k}_raw.csv
done
# SASS of the hot kernels from the built library (what the GPU box ran): mnemonic histogram + full listing of each
python tools/sass_excerpts.py ropebwt3_b200/librb3b200.so profiles/${R}_sass
ls profiles | grep "^${R}_" | wc -l
# roofline.traffic of bench.py: DRAM bytes of the dominant kernel from the ncu --set full capture of this round (static, labelled so)
python - "$R" <<'PY'
import csv, json, sys
R = sys.argv[1]
rows = list(csv.reader(open('profiles/%s_ncu_k_walk_pair_raw.csv' % R)))
h, u, r = rows[0], rows[1], rows[2]
def val(k):
    i = h.index(k); v = float(r[i].replace(',', '')); return v * {'Gbyte': 1e9, 'Mbyte': 1e6, 'Kbyte': 1e3, 'byte': 1}[u[i]]
t = val('dram__bytes_read.sum') + val('dram__bytes_write.sum')
json.dump({"kernel": "k_walk_pair<false>", "dram_bytes_per_launch": t, "rows_per_launch": 10000002,
           "source": "profiles/%s_ncu_k_walk_pair_raw.csv" % R,
           "what": "ncu --set full --clock-control none on bench.py --steps 40 --warmup 3, 31st launch of the kernel (index of ~0.35 G symbols = 0.35 GB of bitmap cells)"},
          open('profiles/walk_first_traffic.json', 'w'))
print("walk kernel DRAM bytes per launch: %.3g" % t)
PY
