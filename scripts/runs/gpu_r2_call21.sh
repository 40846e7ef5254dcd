#!/usr/bin/env bash
# round 2, GPU call 21: MIS rays that cannot reach an emitter are not traced (A/B B200PT_CULL_MIS=0/1 on every scene), full GPU parity suite with it on
set -u
cd /root/repo
mkdir -p gpurun_out
{
for c in 0 1; do
  echo "== CULL_MIS=$c"
  B200PT_CULL_MIS=$c timeout 200 python scripts/perf.py --scene cornell --size 1024 --spp 64 --reps 3 --tag "c2 cull=$c"
  B200PT_CULL_MIS=$c timeout 200 python scripts/perf.py --scene cornell4 --size 256 --spp 256 --reps 3 --tag "c1 cull=$c"
  B200PT_CULL_MIS=$c timeout 200 python scripts/perf.py --scene veach --size 768 --spp 32 --reps 3 --tag "c3 cull=$c"
  B200PT_CULL_MIS=$c timeout 200 python scripts/perf.py --scene vol --size 512 --spp 64 --reps 3 --tag "c5 cull=$c"
  B200PT_CULL_MIS=$c timeout 200 python scripts/perf.py --scene hair --size 512 --spp 32 --reps 3 --tag "hair cull=$c"
  B200PT_CULL_MIS=$c timeout 200 python scripts/perf.py --scene zoo --size 512 --spp 32 --reps 3 --tag "zoo cull=$c"
  B200PT_CULL_MIS=$c timeout 200 python scripts/perf.py --scene zoovpt --size 512 --spp 32 --reps 3 --tag "zoovpt cull=$c"
  B200PT_CULL_MIS=$c timeout 200 python scripts/perf.py --scene smoke --size 1024 --spp 8 --reps 3 --tag "smoke cull=$c"
  B200PT_CULL_MIS=$c timeout 200 python scripts/perf.py --scene shipped --size 1024 --spp 8 --reps 3 --tag "shipped cull=$c"
done
} 2>&1 | grep -E "==|PERF|rror" > gpurun_out/r02u_cull_mis.txt
cat gpurun_out/r02u_cull_mis.txt
timeout 1800 python -m pytest tests/test_gpu_parity.py -m gpu -x -q 2>&1 | tail -8 > gpurun_out/r02u_pytest_gpu.txt
cat gpurun_out/r02u_pytest_gpu.txt
