#!/bin/bash
# 2-GPU session: full gpu suite (incl. the torchrun exchange test), then same-box A/B of the exchanges
N=${1:-2}
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -q -m gpu -x > gpurun_out/t_gpu_all.log 2>&1; echo "pytest rc=$?"; tail -4 gpurun_out/t_gpu_all.log
run() { # rows exchange tag
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 40 --warmup 3 --rows $1 --exchange $2 --no-cpu-baseline > gpurun_out/bench_n${N}_$3.json 2> gpurun_out/bench_n${N}_$3.err; echo "rc=$?"; tail -2 gpurun_out/bench_n${N}_$3.err
python - <<PY
import json
j=json.load(open("gpurun_out/bench_n${N}_$3.json")); r=j["roofline"]; c=j["clocks"]
print("N=$N rows $1 $2: ms/step",round(j["ms_per_step"],3),"q/s",round(j["value"]),"e2e",round(j["e2e"]["value"]),"score ms/step",round(r["score_kernel_share_of_step"]*j["ms_per_step"],3),"sel ms",round(r["select_kernels_ms_per_step"],3),"launches",j["gpu_launches"]/j["steps"],j["config"]["exchange"],"clk",c)
PY
}
for round in 1 2; do
run $((4829565*N)) nccl nccl_4p8_$round
run $((4829565*N)) peer peer_4p8_$round
done
run 38636520 peer peer_full
