#!/bin/bash
mkdir -p gpurun_out
{
for sb in 20480 45056; do
  echo "STAGE_BYTES=$sb"
  B200PT_STAGE_BYTES=$sb timeout 300 python scripts/perf.py --scene veach --size 768 --spp 16 --tag veach_stage$sb
done
for sb in 20480 65536 98304; do
  B200PT_STAGE_BYTES=$sb timeout 300 python scripts/perf.py --scene tris200000 --size 1024 --spp 4 --reps 3 --tag tris200k_stage$sb
done
} 2>&1 | grep -E "PERF|STAGE|rror" | tee gpurun_out/exp.log
