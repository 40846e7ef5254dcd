#!/usr/bin/env bash
# round 2, GPU call 29: ray queue sorted by (any-hit, origin cell, direction octant) before the tree kernel — A/B B200PT_SORT_RAYS=0/1
set -u
cd /root/repo
mkdir -p gpurun_out
{
for s in 0 1; do
  echo "== SORT_RAYS=$s"
  B200PT_SORT_RAYS=$s timeout 300 python scripts/perf.py --scene tris1000000 --size 2048 --spp 4 --reps 2 --tag "c4-1M sort=$s"
  B200PT_SORT_RAYS=$s timeout 300 python scripts/perf.py --scene tris200000 --size 1024 --spp 16 --reps 2 --tag "c4-200k sort=$s"
  B200PT_SORT_RAYS=$s timeout 200 python scripts/perf.py --scene veach --size 768 --spp 32 --reps 3 --tag "c3 sort=$s"
  B200PT_SORT_RAYS=$s timeout 200 python scripts/perf.py --scene hair --size 512 --spp 32 --reps 3 --tag "hair sort=$s"
done
} 2>&1 | grep -E "==|PERF|rror" > gpurun_out/r03c_sort_rays.txt
cat gpurun_out/r03c_sort_rays.txt
B200PT_SORT_RAYS=1 timeout 600 compute-sanitizer --tool memcheck python scripts/compare_ref.py --scene tris20000 --size 96 --spp 2 --no-ref --no-warm 2>&1 | grep -E "ERROR SUMMARY|Invalid" | head -3
B200PT_SORT_RAYS=1 timeout 600 compute-sanitizer --tool racecheck python scripts/compare_ref.py --scene veach --size 96 --spp 2 --no-ref --no-warm 2>&1 | grep -E "RACECHECK SUMMARY|hazard" | head -3
timeout 900 ncu --metrics smsp__inst_executed.sum,smsp__thread_inst_executed.sum,gpu__time_duration.sum --clock-control none --csv -k regex:k_trace python scripts/compare_ref.py --scene tris1000000 --size 2048 --spp 1 --no-ref --no-warm 2>/dev/null | python -c "
import sys, csv
rows=[r for r in csv.reader(sys.stdin) if len(r)>10]
h=next(i for i,r in enumerate(rows) if r[0]=='ID'); H=rows[h]; iN=H.index('Metric Name'); iV=H.index('Metric Value')
acc={}
for r in rows[h+1:]: acc[r[iN]]=acc.get(r[iN],0.0)+float(r[iV].replace(',',''))
print('k_trace sorted: warp inst', acc.get('smsp__inst_executed.sum'), 'lanes', acc.get('smsp__thread_inst_executed.sum',0)/max(acc.get('smsp__inst_executed.sum',1),1))
"
