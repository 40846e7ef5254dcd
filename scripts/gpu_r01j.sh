#!/usr/bin/env bash
mkdir -p gpurun_out
TAG=${1:-r01j}
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1 | cut -c1-250
{
timeout 300 python scripts/perf.py --scene cornell
B200PT_LANES=1 timeout 300 python scripts/perf.py --scene cornell --tag cornell_1lane
B200PT_LANES=3 timeout 300 python scripts/perf.py --scene cornell --tag cornell_3lane
B200PT_LANES=4 timeout 300 python scripts/perf.py --scene cornell --tag cornell_4lane
timeout 300 python scripts/perf.py --scene cornell --pool 2097152
timeout 300 python scripts/perf.py --scene cornell --pool 524288
timeout 300 python scripts/perf.py --scene vol --size 512
timeout 300 python scripts/perf.py --scene veach --size 768 --spp 16
timeout 300 python scripts/perf.py --scene tris200000 --size 1024 --spp 4 --reps 3
B200PT_LANES=1 timeout 300 python scripts/perf.py --scene tris200000 --size 1024 --spp 4 --reps 3 --tag tris_1lane
timeout 300 python scripts/perf.py --scene cornell4 --size 256 --spp 64 --tag c1
B200PT_LANES=1 timeout 300 python scripts/perf.py --scene cornell4 --size 256 --spp 64 --tag c1_1lane
} 2>&1 | grep -E "PERF|DUAL|rror" | tee gpurun_out/perf_$TAG.log
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu_$TAG.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/pytest_gpu_$TAG.log
