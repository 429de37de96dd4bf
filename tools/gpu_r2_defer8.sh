#!/bin/bash
N=${1:-8}
mkdir -p gpurun_out
run() { # tag opts...
tag=$1; shift 1
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 40 --warmup 3 --no-cpu-baseline --no-check "$@" > gpurun_out/df_n${N}_$tag.json 2> gpurun_out/df_n${N}_$tag.err; echo "rc=$?"
python - <<PY
import json
j=json.load(open("gpurun_out/df_n${N}_$tag.json")); r=j["roofline"]
print("N=$N $tag: ms/step",round(j["ms_per_step"],4),"q/s",round(j["value"]),"e2e",round(j["e2e"]["value"]),"per_rank",j["per_rank"])
PY
}
run defer1_a
run defer0_a --opt xchg_defer=0
run defer1_b
run defer0_b --opt xchg_defer=0
