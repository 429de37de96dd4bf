#!/bin/bash
# ncu --set full of the small kernels of a TS pass (finalize, bootstrap refresh, dense launch)
mkdir -p gpurun_out
timeout 1200 ncu --set full --clock-control none --import-source on -k regex:"finalize_kernel|refresh_kernel" -s 6 -c 2 -o gpurun_out/prof_small python bench.py --rows 4829565 --steps 2 --warmup 3 --no-cpu-baseline --no-check > gpurun_out/ncu_small.log 2>&1; echo "ncu rc=$?"; tail -2 gpurun_out/ncu_small.log
timeout 1200 ncu --set full --clock-control none --import-source on -k regex:umma_score -s 6 -c 1 -o gpurun_out/prof_dense python bench.py --rows 4829565 --steps 2 --warmup 3 --no-cpu-baseline --no-check > gpurun_out/ncu_dense.log 2>&1; echo "ncu rc=$?"; tail -2 gpurun_out/ncu_dense.log
