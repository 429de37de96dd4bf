#!/bin/bash
mkdir -p gpurun_out
echo "=== pytest tensor engines"
timeout 900 python -m pytest tests/test_gpu_parity.py -q -m gpu -k "umma" --maxfail=6 > gpurun_out/t_umma.log 2>&1; echo "rc=$?"; tail -6 gpurun_out/t_umma.log
for cfg in "1 1 4829565 20" "1 0 4829565 20" "2 1 4829565 20" "1 1 38636520 30" "1 0 38636520 30" "2 1 38636520 30"; do
set -- $cfg
timeout 900 python bench.py --rows $3 --steps $4 --variant $1 --l2-prefetch $2 --no-cpu-baseline --no-check > gpurun_out/bench_v$1_p$2_$3.json 2> gpurun_out/bench_v$1_p$2_$3.err; echo "rc=$?"; tail -2 gpurun_out/bench_v$1_p$2_$3.err
python - <<PY
import json
j=json.load(open("gpurun_out/bench_v$1_p$2_$3.json")); r=j["roofline"]; c=j["clocks"]
print("variant $1 prefetch $2 rows $3: ms/step",round(j["ms_per_step"],3),"q/s",round(j["value"]),"kernel GB/s",round(r["achieved"]),"ms/launch",round(r["ms_per_launch"],3),"sel ms",round(r["select_kernels_ms_per_step"],3),"clk",c.get("sm_mhz"),c.get("sm_mhz_min"),c.get("reasons"),c.get("power_w_median"))
PY
done
