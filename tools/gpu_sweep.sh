#!/bin/bash
mkdir -p gpurun_out
timeout 1200 python tools/sweep.py 8841823 100 gpurun_out/sweep_8p8M_k100.json 2>&1 | tail -30
timeout 900 python tools/sweep.py 8841823 1000 gpurun_out/sweep_8p8M_k1000.json 8,173,1024 2>&1 | tail -8
