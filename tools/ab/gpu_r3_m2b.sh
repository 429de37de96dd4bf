#!/bin/bash
# quick 2-GPU validation of the direct-to-host result path (b2f_search, dist.search_host)
mkdir -p gpurun_out
timeout 400 python -m pytest tests -q -m gpu --timeout 300 -x -k "peer_memory or multi_gpu or config1 or golden or resident or device_resident or asynchronous or overflowed or identical_rows or empty_index" > gpurun_out/r3m2b_tests.log 2>&1; echo "rc=$?"; tail -4 gpurun_out/r3m2b_tests.log
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 20 --warmup 5 --rows 9659130 > gpurun_out/r3m2b_bench.json 2> gpurun_out/r3m2b_bench.err; echo "rc=$?"
python - <<'PY'
import json
try:
    j=json.load(open("gpurun_out/r3m2b_bench.json")); r=j["roofline"]
    print("ms/step",round(j["ms_per_step"],3),"q/s",round(j["value"]),"e2e",round(j["e2e"]["value"]),round(j["e2e"]["ms_per_step"],3),"frac",round(r["frac"],3),"sus",round(j["sustained"]["ms_per_step"],3),"inproc",j.get("inproc"),"check",j["check"])
except Exception as e:
    print("FAILED",e); print(open("gpurun_out/r3m2b_bench.err").read()[-2000:])
PY
timeout 200 python bench.py --rows 4829565 --steps 40 --no-cpu-baseline --sustain-seconds 0 > gpurun_out/r3m2b_b1.json 2> gpurun_out/r3m2b_b1.err; echo "rc=$?"
python - <<'PY'
import json
try:
    j=json.load(open("gpurun_out/r3m2b_b1.json")); r=j["roofline"]
    print("N=1 4.8M ms/step",round(j["ms_per_step"],3),"q/s",round(j["value"]),"e2e",round(j["e2e"]["value"]),round(j["e2e"]["ms_per_step"],3),"frac",round(r["frac"],3),"check",j["check"])
except Exception as e:
    print("FAILED",e); print(open("gpurun_out/r3m2b_b1.err").read()[-2000:])
PY
