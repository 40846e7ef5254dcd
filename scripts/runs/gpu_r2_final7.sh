#!/usr/bin/env bash
# round 2, last GPU seconds: smoke() on the final tree (the Python scene loader changed after the last full GPU suite)
set -u
cd /root/repo
mkdir -p gpurun_out
( time timeout 30 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" ) 2>&1 | tail -6 > gpurun_out/r10_smoke.txt
cat gpurun_out/r10_smoke.txt
