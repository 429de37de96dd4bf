#!/bin/bash
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_parity.py -q -m gpu -x -k "merge or multi_gpu or peer_memory" > gpurun_out/t_last.log 2>&1; echo "pytest rc=$?"; tail -2 gpurun_out/t_last.log
timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 20 --warmup 3 --rows 9659130 --no-cpu-baseline > gpurun_out/last_n2.json 2> gpurun_out/last_n2.err; echo "bench rc=$?"
python - <<PY
import json
j=json.load(open("gpurun_out/last_n2.json"))
print("N=2: ms/step",round(j["ms_per_step"],4),"q/s",round(j["value"]),"e2e",round(j["e2e"]["value"]),j["check"],j["per_rank"])
PY
grep -c "terminate called" gpurun_out/last_n2.err
