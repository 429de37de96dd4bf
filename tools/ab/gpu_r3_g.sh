#!/bin/bash
# 2-GPU validation of two-round rescoring + the one-call collective host path, then the aniso comparison
mkdir -p gpurun_out
timeout 500 python -m pytest tests -q -m gpu --timeout 300 -x > gpurun_out/r3g_tests.log 2>&1; echo "rc=$?"; tail -4 gpurun_out/r3g_tests.log
b2() { tag=$1; shift
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 20 --warmup 5 --rows 9659130 --inproc 0 --no-oracle-check "$@" > gpurun_out/r3g_$tag.json 2> gpurun_out/r3g_$tag.err; echo "rc=$?"
python - <<PY
import json
try:
    j=json.load(open("gpurun_out/r3g_$tag.json")); r=j["roofline"]
    print("$tag ms/step",round(j["ms_per_step"],3),"q/s",round(j["value"]),"e2e",round(j["e2e"]["value"]),round(j["e2e"]["ms_per_step"],3),"frac",round(r["frac"],3),"ms/launch",round(r["ms_per_launch"],3),"sel",round(r["select_kernels_ms_per_step"],3),"sus",round(j["sustained"]["ms_per_step"],3),"per_rank",j["per_rank"],"fb",j["check"]["fallback_queries"])
except Exception as e:
    print("$tag FAILED",e); print(open("gpurun_out/r3g_$tag.err").read()[-2000:])
PY
}
b2 n2_iso
b2 n2_aniso --data aniso
b1() { tag=$1; shift
timeout 200 python bench.py --rows 4829565 --steps 40 --no-cpu-baseline --sustain-seconds 0 --no-oracle-check "$@" > gpurun_out/r3g_$tag.json 2> gpurun_out/r3g_$tag.err; echo "rc=$?"
python - <<PY
import json
try:
    j=json.load(open("gpurun_out/r3g_$tag.json")); r=j["roofline"]
    print("$tag ms/step",round(j["ms_per_step"],3),"q/s",round(j["value"]),"e2e",round(j["e2e"]["value"]),round(j["e2e"]["ms_per_step"],3),"frac",round(r["frac"],3),"ms/launch",round(r["ms_per_launch"],3),"sel",round(r["select_kernels_ms_per_step"],3),"fb",j["check"]["fallback_queries"])
except Exception as e:
    print("$tag FAILED",e); print(open("gpurun_out/r3g_$tag.err").read()[-2000:])
PY
}
b1 n1_iso
b1 n1_aniso --data aniso
b1 n1_k1000 --k 1000 --steps 20
b1 n1_ts_iso --variant 2
