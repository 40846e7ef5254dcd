#!/usr/bin/env bash
mkdir -p gpurun_out
TAG=${1:-r01}
N=${2:-1}
if [ "$N" = "1" ]; then
  python bench.py --workload c4 --steps 3 --warmup 1 --spp-per-step 32 --no-cpu-baseline > gpurun_out/bench_c4_$TAG.json 2> gpurun_out/bench_c4_$TAG.err; tail -1 gpurun_out/bench_c4_$TAG.json | cut -c1-600; tail -2 gpurun_out/bench_c4_$TAG.err
else
  python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29519 bench.py --gpus $N --workload c4 --steps 3 --warmup 1 --spp-per-step 32 > gpurun_out/bench_c4_${TAG}_n$N.json 2> gpurun_out/bench_c4_${TAG}_n$N.err
  tail -1 gpurun_out/bench_c4_${TAG}_n$N.json | cut -c1-600; tail -2 gpurun_out/bench_c4_${TAG}_n$N.err
fi
