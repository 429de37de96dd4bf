#!/bin/bash
mkdir -p gpurun_out
timeout 500 python -m pytest tests -q -m gpu --timeout 300 > gpurun_out/r3h_tests.log 2>&1; echo "rc=$?"; tail -4 gpurun_out/r3h_tests.log
b1() { tag=$1; shift
timeout 200 python bench.py --rows 4829565 --steps 40 --no-cpu-baseline --sustain-seconds 0 "$@" > gpurun_out/r3h_$tag.json 2> gpurun_out/r3h_$tag.err; echo "rc=$?"
python - <<PY
import json
try:
    j=json.load(open("gpurun_out/r3h_$tag.json")); r=j["roofline"]; c=j["check"]
    print("$tag ms/step",round(j["ms_per_step"],3),"q/s",round(j["value"]),"e2e",round(j["e2e"]["value"]),round(j["e2e"]["ms_per_step"],3),"frac",round(r["frac"],3),"ms/launch",round(r["ms_per_launch"],3),"sel",round(r["select_kernels_ms_per_step"],3),"fb",c["fallback_queries"],"oracle",c.get("oracle_violations"),c.get("oracle_exact_rows"),"redo",c.get("overflow_redo_ok"))
except Exception as e:
    print("$tag FAILED",e); print(open("gpurun_out/r3h_$tag.err").read()[-2000:])
PY
}
b1 n1_iso
b1 n1_aniso --data aniso
b1 n1_k1000 --k 1000 --steps 20
b1 n1_aniso_k1000 --k 1000 --steps 20 --data aniso
