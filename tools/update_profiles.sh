#!/bin/bash
# Copies the judged summaries from gpurun_out/ (scratch) into profiles/ (tracked).
set -e
cd "$(dirname "$0")/.."
cp gpurun_out/launches.csv profiles/r1_launches_bench_steps10.csv
cp gpurun_out/bench_full.json profiles/r1_bench_full.json
cp gpurun_out/bench_ref.json profiles/r1_bench_reference_arm.json
cp gpurun_out/bench_g10.json profiles/r1_bench_10_genomes_per_merge.json
cp gpurun_out/rank_bench_bm.jsonl profiles/r1_rank_bench_bitmap.jsonl
cp gpurun_out/rank_bench_rle.jsonl profiles/r1_rank_bench_rle.jsonl
python tools/ncu_summary.py gpurun_out/prof_walk_bm.ncu-rep > profiles/r1_ncu_walk_first_bm.txt 2>/dev/null
python tools/ncu_summary.py gpurun_out/prof_lf_bm.ncu-rep > profiles/r1_ncu_lf_bm.txt 2>/dev/null
python tools/ncu_summary.py gpurun_out/prof_fix_bm.ncu-rep > profiles/r1_ncu_walk_fix_log.txt 2>/dev/null
ncu -i gpurun_out/prof_fix_bm.ncu-rep --page raw --csv 2>/dev/null > profiles/r1_ncu_walk_fix_log_raw.csv
ncu -i gpurun_out/prof_walk_bm.ncu-rep --page raw --csv 2>/dev/null > profiles/r1_ncu_walk_first_bm_raw.csv
ncu -i gpurun_out/prof_lf_bm.ncu-rep --page raw --csv 2>/dev/null > profiles/r1_ncu_lf_bm_raw.csv
python - <<'PY'
import csv, json
rows=list(csv.reader(open('profiles/r1_ncu_walk_first_bm_raw.csv')))
h,u,r=rows[0],rows[1],rows[2]
def val(k):
    i=h.index(k); v=float(r[i].replace(',','')); return v*{'Gbyte':1e9,'Mbyte':1e6,'Kbyte':1e3,'byte':1}[u[i]]
t=val('dram__bytes_read.sum')+val('dram__bytes_write.sum')
json.dump({"kernel":"k_walk_first<BmPair>","dram_bytes_per_launch":t,"rows_per_launch":10000002,"source":"profiles/r1_ncu_walk_first_bm_raw.csv (ncu --set full on bench.py --steps 64 --warmup 3, 61st launch: index of 0.64 G symbols)"}, open('profiles/walk_first_traffic.json','w'))
rows=[x for x in csv.reader(open('profiles/r1_launches_bench_steps10.csv')) if len(x)>5]
hdr=None; agg={}
for x in rows:
    if x[0]=='ID': hdr=x; continue
    if hdr is None: continue
    n=x[hdr.index('Kernel Name')].split('(')[0]; v=float(x[hdr.index('Metric Value')].replace(',','')); un=x[hdr.index('Metric Unit')]
    v*= {'ns':1e-3,'us':1,'ms':1e3,'s':1e6}[un]
    a=agg.setdefault(n,[0,0.0]); a[0]+=1; a[1]+=v
tot=sum(a[1] for a in agg.values())
with open('profiles/r1_launch_shares.txt','w') as f:
    f.write("ncu --metrics gpu__time_duration.sum --clock-control none -k regex:k_|Kernel ... python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-e2e --no-rank-bench\n")
    f.write("(cold-cache, serialised launches: compare shares, not absolutes; 14 genomes = 14 BWT builds, 1 index build, 13 merges)\n")
    for n,a in sorted(agg.items(), key=lambda kv:-kv[1][1]):
        f.write("%-44s n=%4d total=%10.1f us  %5.1f%%\n"%(n[:44],a[0],a[1],100*a[1]/tot))
print(open('profiles/r1_launch_shares.txt').read())
PY
