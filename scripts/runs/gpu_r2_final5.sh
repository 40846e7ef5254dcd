#!/usr/bin/env bash
# round 2, last GPU seconds: ncu counters of the CTA-local lambertian kernel on the FINAL binary (c2, c1).  The counters
# bench.py reads for these two workloads were taken before the kernel went from 3 x 256 to 2 x 448 slots per SM (d69c70b).
set -u
cd /root/repo
mkdir -p gpurun_out
cp profiles/r02_counters.json gpurun_out/r10_counters.json
timeout 55 python scripts/ncu_counters.py --out gpurun_out/r10_counters.json --workloads c2 2>&1 | grep -v "^    " | cut -c1-400 > gpurun_out/r10_counters.log
timeout 35 python scripts/ncu_counters.py --out gpurun_out/r10_counters.json --workloads c1 2>&1 | grep -v "^    " | cut -c1-400 >> gpurun_out/r10_counters.log
cat gpurun_out/r10_counters.log
