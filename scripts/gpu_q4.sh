#!/usr/bin/env bash
for sb in 0 8192 20480 49152; do
  B200PT_STAGE_BYTES=$sb timeout 300 python scripts/perf.py --scene veach --size 768 --spp 32 --reps 3 --tag "stage=$sb veach" 2>&1 | grep PERF | cut -c1-110
  B200PT_STAGE_BYTES=$sb timeout 300 python scripts/perf.py --scene tris200000 --size 1024 --spp 4 --reps 3 --tag "stage=$sb tris200k" 2>&1 | grep PERF | cut -c1-110
done
