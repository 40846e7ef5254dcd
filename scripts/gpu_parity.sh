#!/usr/bin/env bash
mkdir -p gpurun_out
TAG=${1:-p}
for cfg in "cornell 1024 32" "veach 768 64" "vol 512 64" "tris200000 512 16"; do
  set -- $cfg
  timeout 600 python scripts/compare_ref.py --scene $1 --size $2 --spp $3 2>&1 | grep parity | sed "s/^/$1 $3spp /" | cut -c1-200
done | tee gpurun_out/parity_$TAG.log
