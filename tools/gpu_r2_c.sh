#!/bin/bash
# SS variant: L2 prefetch distance sweep against TS, short (4.8M) and long (19.3M / 38.6M) steps
mkdir -p gpurun_out
run() { # rows steps variant pf tighten tag
timeout 900 python bench.py --rows $1 --steps $2 --variant $3 --l2-prefetch $4 --tighten $5 --no-cpu-baseline --no-check > gpurun_out/ss_$6.json 2> gpurun_out/ss_$6.err; tail -2 gpurun_out/ss_$6.err
python - <<PY
import json
j=json.load(open("gpurun_out/ss_$6.json")); r=j["roofline"]; c=j["clocks"]
print("$6 rows $1 variant $3 pf $4: ms/step",round(j["ms_per_step"],3),"q/s",round(j["value"]),"score ms/step",round(r["score_kernel_share_of_step"]*j["ms_per_step"],3),"sel ms",round(r["select_kernels_ms_per_step"],3),"launches/step",j["gpu_launches"]/j["steps"],"clk",c.get("sm_mhz"),c.get("sm_mhz_min"),c.get("power_w_median"),c.get("reasons"))
PY
}
run 4829565 40 2 1 400 ts_4p8
run 4829565 40 1 1 0 ss1_4p8
run 4829565 40 1 2 0 ss2_4p8
run 4829565 40 1 3 0 ss3_4p8
run 4829565 40 1 0 0 ss0_4p8
run 19318260 30 2 1 400 ts_19
run 19318260 30 1 1 0 ss1_19
run 19318260 30 1 2 0 ss2_19
run 19318260 30 1 3 0 ss3_19
run 38636520 20 2 1 400 ts_38
run 38636520 20 1 2 0 ss2_38
run 38636520 20 1 3 0 ss3_38
