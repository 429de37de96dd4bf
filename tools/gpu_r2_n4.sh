#!/bin/bash
# N-GPU default bench only (what the driver's scaling run launches)
N=${1:-4}
mkdir -p gpurun_out
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 30 --warmup 3 > gpurun_out/bench_n${N}.json 2> gpurun_out/bench_n${N}.err; echo "rc=$?"; tail -2 gpurun_out/bench_n${N}.err
cat gpurun_out/bench_n${N}.json
