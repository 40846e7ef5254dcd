#!/usr/bin/env bash
for L in 2 3 4; do
  for cfg in "cornell 1024 32" "veach 768 32" "vol 512 32" "tris200000 1024 4"; do
    set -- $cfg
    B200PT_LANES=$L timeout 300 python scripts/perf.py --scene $1 --size $2 --spp $3 --reps 3 --tag "lanes=$L $1" 2>&1 | grep PERF | cut -c1-110
  done
done
