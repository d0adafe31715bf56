#!/bin/bash
# does cudaLimitMaxL2FetchGranularity change what the bitmap rank kernel pulls from DRAM?  (64-B half cell needed per query)
O=gpurun_out
for G in 32 64 128; do
  RB3B_L2_FETCH=$G timeout 200 python tools/rank_bench.py --kind bitmap --reps 5 2>/dev/null | python -c "
import sys, json
for l in sys.stdin:
    d = json.loads(l); print('L2_FETCH=$G no-profiler: %.2f G queries/s, frac %.3f' % (d['gqueries_per_s'], d['frac']))"
  RB3B_L2_FETCH=$G timeout 300 ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --clock-control none -k regex:k_lf_bm -s 2 -c 1 --csv --log-file $O/r2_l2fetch_$G.csv python tools/rank_bench.py --kind bitmap --reps 1 > /dev/null 2>&1
  grep -E "dram__bytes|gpu__time" $O/r2_l2fetch_$G.csv | awk -F'","' '{print "   L2_FETCH='$G' ncu:", $(NF-2), $(NF-1), $NF}'
done
