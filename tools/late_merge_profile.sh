timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"k_" -s 2400 -c 400 --csv --log-file gpurun_out/launches_late.csv python bench.py --steps 66 --warmup 3 --no-cpu-baseline --no-e2e --no-rank-bench > /dev/null 2>&1
python - <<'PY'
import csv
rows=[x for x in csv.reader(open('gpurun_out/launches_late.csv')) if len(x)>5]
hdr=None; seq=[]
for x in rows:
    if x[0]=='ID': hdr=x; continue
    if hdr is None: continue
    n=x[hdr.index('Kernel Name')].split('(')[0]; v=float(x[hdr.index('Metric Value')].replace(',','')); un=x[hdr.index('Metric Unit')]
    v*= {'ns':1e-3,'us':1,'ms':1e3,'s':1e6}[un]
    seq.append((n.replace('void ',''),v))
# last merge
idx=[i for i,(n,v) in enumerate(seq) if n.startswith('k_prep_count')]
a=idx[-2]; b=idx[-1]
tot=0
for n,v in seq[a:b]:
    tot+=v
agg={}
for n,v in seq[a:b]:
    agg[n]=agg.get(n,0)+v
for n,v in sorted(agg.items(), key=lambda kv:-kv[1]): print("%-28s %8.1f us"%(n[:28],v))
print("sum", tot)
PY
