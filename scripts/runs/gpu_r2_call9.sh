#!/usr/bin/env bash
# round 2, GPU call 9 (2 GPUs): sorted trace phase A/B + sanitizer, het CTA size A/B, multi-GPU tests (NCCL inside the library), bench at N = 2 under torchrun
set -u
mkdir -p gpurun_out
{
for lib in gpu-pathtracer_b200/csrc/libb200pt.so gpu-pathtracer_b200/csrc/libb200pt_nosort.so; do
  timeout 200 python scripts/perf.py --scene cornell --size 1024 --spp 64 --reps 5 --lib $lib --tag "c2 $(basename $lib)"
  timeout 200 python scripts/perf.py --scene cornell4 --size 256 --spp 64 --reps 5 --lib $lib --tag "c1 $(basename $lib)"
  timeout 200 python scripts/perf.py --scene vol --size 512 --spp 64 --reps 5 --lib $lib --tag "c5 $(basename $lib)"
  timeout 200 python scripts/perf.py --scene zoo --size 512 --spp 32 --reps 3 --lib $lib --tag "zoo $(basename $lib)"
done
for lib in gpu-pathtracer_b200/csrc/libb200pt.so gpu-pathtracer_b200/csrc/libb200pt_het256.so; do
  timeout 200 python scripts/perf.py --scene smoke --size 1024 --spp 16 --reps 3 --lib $lib --tag "smoke $(basename $lib)"
  timeout 200 python scripts/perf.py --scene shipped --size 512 --spp 16 --reps 3 --lib $lib --tag "shipped $(basename $lib)"
done
timeout 600 compute-sanitizer --tool racecheck --error-exitcode 9 python scripts/perf.py --scene cornell --size 64 --spp 1 --reps 1 2>&1 | tail -3
timeout 600 compute-sanitizer --tool memcheck --error-exitcode 9 python scripts/perf.py --scene smoke --size 64 --spp 1 --reps 1 2>&1 | tail -3
timeout 600 compute-sanitizer --tool racecheck --error-exitcode 9 python scripts/perf.py --scene smoke --size 64 --spp 1 --reps 1 2>&1 | tail -3
} 2>&1 | grep -E "PERF|SUMMARY|rror" > gpurun_out/r02i_sort_ab.txt
( timeout 900 python -m pytest tests/test_gpu_multi.py -m gpu -q 2>&1 | tail -8 ) > gpurun_out/r02i_pytest_multi.txt
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus 2 --steps 8 --warmup 3 > gpurun_out/r02i_bench_c2_n2.json 2> gpurun_out/r02i_bench_c2_n2.err
cat gpurun_out/r02i_sort_ab.txt; cat gpurun_out/r02i_pytest_multi.txt; cut -c1-600 gpurun_out/r02i_bench_c2_n2.json; tail -5 gpurun_out/r02i_bench_c2_n2.err
