#!/bin/bash
# usage: tools/ncu_one.sh <kernel regex> <tag> [skip] [bench args...]: one ncu --set full capture of one launch inside a bench run
k=$1; tag=$2; skip=${3:-30}; shift 3
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"$k" -s $skip -c 1 -f -o gpurun_out/prof_$tag python bench.py --steps 40 --warmup 3 --no-cpu-baseline --no-e2e --no-rank-bench "$@" > gpurun_out/ncu_$tag.log 2>&1
python tools/ncu_summary.py gpurun_out/prof_$tag.ncu-rep > gpurun_out/ncu_$tag.txt 2>&1
cat gpurun_out/ncu_$tag.txt
