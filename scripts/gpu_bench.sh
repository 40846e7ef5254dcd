#!/usr/bin/env bash
# official bench lines (ours + reference arm) on N GPUs, plus the launch list / DRAM-bytes capture on 1 GPU
mkdir -p gpurun_out
TAG=${1:-r01}
N=${2:-1}
if [ "$N" = "1" ]; then
  python bench.py --steps 8 --warmup 3 > gpurun_out/bench_$TAG.json 2> gpurun_out/bench_$TAG.err; tail -1 gpurun_out/bench_$TAG.json; tail -3 gpurun_out/bench_$TAG.err
  python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_ref_$TAG.json 2> gpurun_out/bench_ref_$TAG.err; tail -1 gpurun_out/bench_ref_$TAG.json; tail -3 gpurun_out/bench_ref_$TAG.err
  ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -c 3000 --csv --log-file gpurun_out/launches_$TAG.csv \
      python bench.py --steps 1 --warmup 0 --spp-per-step 4 --no-cpu-baseline > gpurun_out/ncu_bench_$TAG.log 2>&1
  tail -2 gpurun_out/ncu_bench_$TAG.log | cut -c1-300
else
  python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus $N --steps 8 --warmup 3 > gpurun_out/bench_${TAG}_n$N.json 2> gpurun_out/bench_${TAG}_n$N.err
  tail -1 gpurun_out/bench_${TAG}_n$N.json; tail -3 gpurun_out/bench_${TAG}_n$N.err
fi
