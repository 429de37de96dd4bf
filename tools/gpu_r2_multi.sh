#!/bin/bash
# multi-GPU session: in-process sharded facade test + torchrun benches at N=$1
N=${1:-2}
mkdir -p gpurun_out
nvidia-smi -L | head -8
timeout 600 python -m pytest tests/test_gpu_parity.py -q -m gpu -k "multi_gpu" > gpurun_out/t_multi.log 2>&1; echo "multi_gpu test rc=$?"; tail -3 gpurun_out/t_multi.log
run() { # rows tag
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 30 --warmup 3 --rows $1 > gpurun_out/bench_n${N}_$2.json 2> gpurun_out/bench_n${N}_$2.err; echo "rc=$?"; tail -3 gpurun_out/bench_n${N}_$2.err
python - <<PY
import json
j=json.load(open("gpurun_out/bench_n${N}_$2.json")); r=j["roofline"]; c=j["clocks"]
print("N=$N rows $1: ms/step",round(j["ms_per_step"],3),"q/s",round(j["value"]),"e2e",round(j["e2e"]["value"]),"kernel GB/s",round(r["achieved"]),"score ms/step",round(r["score_kernel_share_of_step"]*j["ms_per_step"],3),"sel ms",round(r["select_kernels_ms_per_step"],3),"launches",j["gpu_launches"]/j["steps"],"clk",c.get("sm_mhz"),c.get("reasons"),"check",j["check"])
PY
}
run 38636520 full
run $((4829565*N)) pergpu4p8M
