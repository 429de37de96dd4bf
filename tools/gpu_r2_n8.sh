#!/bin/bash
# N-GPU default bench (what the driver's scaling run launches) + reference arm
N=${1:-8}
mkdir -p gpurun_out
nvidia-smi -L | wc -l
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 30 --warmup 3 > gpurun_out/bench_n${N}.json 2> gpurun_out/bench_n${N}.err; echo "rc=$?"; tail -3 gpurun_out/bench_n${N}.err
cat gpurun_out/bench_n${N}.json
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29512 bench.py --impl reference --gpus $N --steps 3 --warmup 1 > gpurun_out/bench_ref_n${N}.json 2> gpurun_out/bench_ref_n${N}.err; echo "rc=$?"; cat gpurun_out/bench_ref_n${N}.json
