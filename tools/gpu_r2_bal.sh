#!/bin/bash
# same-box A/B at N GPUs: speed-weighted shards vs equal shards
N=${1:-8}
mkdir -p gpurun_out
run() { # balance tag
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 40 --warmup 3 --balance $1 --no-cpu-baseline --no-check > gpurun_out/bal_n${N}_$2.json 2> gpurun_out/bal_n${N}_$2.err; echo "rc=$?"
python - <<PY
import json
j=json.load(open("gpurun_out/bal_n${N}_$2.json")); r=j["roofline"]
print("N=$N balance=$1 $2: ms/step",round(j["ms_per_step"],4),"q/s",round(j["value"]),"e2e",round(j["e2e"]["value"]),"rank0 score ms/step",round(r["score_kernel_share_of_step"]*j["ms_per_step"],3),j["config"]["balance"])
PY
}
run 0 eq1
run 1 bal1
run 0 eq2
run 1 bal2
