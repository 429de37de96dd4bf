#!/bin/bash
# multi-GPU session: in-process sharded facade test + torchrun bench at N=$1
N=${1:-2}
mkdir -p gpurun_out
nvidia-smi -L
echo "=== in-process multi-GPU facade test"
timeout 600 python -m pytest tests/test_gpu_parity.py -q -m gpu -k "multi_gpu" > gpurun_out/t_multi.log 2>&1; echo "rc=$?"; tail -5 gpurun_out/t_multi.log
echo "=== torchrun bench N=$N"
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 20 --warmup 3 > gpurun_out/bench_n$N.json 2> gpurun_out/bench_n$N.err; echo "rc=$?"; tail -5 gpurun_out/bench_n$N.err; cat gpurun_out/bench_n$N.json
echo "=== reference arm under torchrun"
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29512 bench.py --impl reference --gpus $N --steps 3 --warmup 1 > gpurun_out/bench_ref_n$N.json 2> gpurun_out/bench_ref_n$N.err; echo "rc=$?"; tail -3 gpurun_out/bench_ref_n$N.err; cat gpurun_out/bench_ref_n$N.json
