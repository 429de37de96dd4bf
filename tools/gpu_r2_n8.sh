#!/bin/bash
# N-GPU session: multi-GPU tests, default bench under torchrun (what the driver's scaling run launches),
# NCCL fallback for comparison, reference arm, in-process sharded index
N=${1:-8}
mkdir -p gpurun_out
nvidia-smi -L | wc -l
timeout 900 python -m pytest tests/test_gpu_parity.py -q -m gpu -k "multi_gpu or peer_memory" > gpurun_out/t_multi.log 2>&1; echo "multi-gpu tests rc=$?"; tail -3 gpurun_out/t_multi.log
run() { # exchange tag extra...
ex=$1; tag=$2; shift 2
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 30 --warmup 3 --exchange $ex "$@" > gpurun_out/bench_n${N}$tag.json 2> gpurun_out/bench_n${N}$tag.err; echo "rc=$?"; tail -2 gpurun_out/bench_n${N}$tag.err
python - <<PY
import json
j=json.load(open("gpurun_out/bench_n${N}$tag.json")); r=j["roofline"]; c=j["clocks"]
print("N=$N $ex $tag: ms/step",round(j["ms_per_step"],3),"q/s",round(j["value"]),"e2e",round(j["e2e"]["value"]),"score ms/step",round(r["score_kernel_share_of_step"]*j["ms_per_step"],3),"sel ms",round(r["select_kernels_ms_per_step"],3),"launches",j["gpu_launches"]/j["steps"],j["config"]["exchange"],"clk",c.get("sm_mhz"),c.get("reasons"),"check",j["check"])
PY
}
run peer ""
run nccl _nccl --no-cpu-baseline
run peer _again --no-cpu-baseline
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29512 bench.py --impl reference --gpus $N --steps 3 --warmup 1 > gpurun_out/bench_ref_n${N}.json 2> gpurun_out/bench_ref_n${N}.err; echo "rc=$?"; cat gpurun_out/bench_ref_n${N}.json
timeout 600 python tools/inproc_bench.py $N 38636520 > gpurun_out/inproc_n${N}.json 2> gpurun_out/inproc_n${N}.err; echo "rc=$?"; tail -2 gpurun_out/inproc_n${N}.err; cat gpurun_out/inproc_n${N}.json
