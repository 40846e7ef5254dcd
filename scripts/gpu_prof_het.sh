#!/usr/bin/env bash
# GPU box: the two smoke parity tests that changed, then one ncu --set full capture of the heterogeneous-media kernel
# and of the reference's Volpath on the same scene.
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_parity.py -x -q -k "cpu_oracle and smoke" 2>&1 | tail -3
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_volpath_seq -c 1 -f -o gpurun_out/prof_het_seq \
    python scripts/compare_ref.py --scene smoke --size 512 --spp 8 --no-ref > gpurun_out/ncu_het_seq.log 2>&1
timeout 600 ncu --set full --clock-control none -k regex:Volpath -s 2 -c 1 -f -o gpurun_out/prof_het_ref \
    python scripts/compare_ref.py --scene smoke --size 512 --spp 4 > gpurun_out/ncu_het_ref.log 2>&1
ls -la gpurun_out | grep het
