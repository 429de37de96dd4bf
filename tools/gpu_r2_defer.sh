#!/bin/bash
# N GPUs: exchange test (N>=2), then same-box A/B of the deferred merge (and of speed-weighted shards)
N=${1:-2}
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py -q -m gpu -k "multi_gpu or peer_memory" > gpurun_out/t_multi.log 2>&1; echo "multi-gpu tests rc=$?"; tail -3 gpurun_out/t_multi.log
run() { # tag rows opts...
tag=$1; rows=$2; shift 2
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 40 --warmup 3 --rows $rows --no-cpu-baseline "$@" > gpurun_out/df_n${N}_$tag.json 2> gpurun_out/df_n${N}_$tag.err; echo "rc=$?"; tail -1 gpurun_out/df_n${N}_$tag.err
python - <<PY
import json
j=json.load(open("gpurun_out/df_n${N}_$tag.json")); r=j["roofline"]
print("N=$N $tag: ms/step",round(j["ms_per_step"],4),"q/s",round(j["value"]),"e2e",round(j["e2e"]["value"]),"ok",j["check"]["host_api_equals_device_api"],j["check"]["tensor_engine_equals_simt_engine_4q"],"per_rank",j["per_rank"])
PY
}
ROWS=$((4829565*N))
run defer1_a $ROWS --balance 0
run defer0_a $ROWS --balance 0 --opt xchg_defer=0
run defer1_b $ROWS --balance 0
run defer0_b $ROWS --balance 0 --opt xchg_defer=0
run defer1_bal $ROWS --balance 1
