#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py -q -m gpu -k "umma" --maxfail=6 > gpurun_out/t_umma.log 2>&1; echo "rc=$?"; tail -3 gpurun_out/t_umma.log
timeout 1200 ncu --set full --clock-control none --import-source on -k regex:umma_ss -s 8 -c 1 -o gpurun_out/prof_ss python bench.py --rows 4829565 --steps 2 --warmup 3 --variant 1 --no-cpu-baseline --no-check > gpurun_out/ncu_full.log 2>&1; echo "rc=$?"; tail -2 gpurun_out/ncu_full.log
