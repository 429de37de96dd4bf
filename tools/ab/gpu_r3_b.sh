#!/bin/bash
# round 2, call B: what bounds the scoring kernels?  ncu --set full of QS (R=6, R=0, R=12) and TS on a 4.83M-row
# shard; sustained (power-capped) behaviour as a function of the batch size; a plain sustained copy.
mkdir -p gpurun_out
run() { # tag, args...
  tag=$1; shift
  timeout 600 python bench.py --no-cpu-baseline --no-check "$@" > gpurun_out/r3b_$tag.json 2> gpurun_out/r3b_$tag.err; rc=$?
  python - <<PY
import json
try:
    j=json.load(open("gpurun_out/r3b_$tag.json")); r=j["roofline"]; c=j["clocks"]
    print("$tag rc=$rc ms/step",round(j["ms_per_step"],3),"q/s",round(j["value"]),"kernel GB/s",round(r["achieved"]),"frac",round(r["frac"],3),"ms/launch",round(r["ms_per_launch"],3),"sel ms",round(r["select_kernels_ms_per_step"],3),r["kernel"][:12],"clk",c.get("sm_mhz"),c.get("sm_mhz_min"),c.get("power_w_median"),c.get("reasons"))
except Exception as e:
    print("$tag rc=$rc FAILED", e); print(open("gpurun_out/r3b_$tag.err").read()[-1500:])
PY
}
echo "=== sustained (>= 1 s) runs at the full size, by batch size"
for nq in 16 64 128 173; do
run qs_38_nq$nq --variant 3 --nq $nq --steps 100
done
run qs6_38_nq173 --variant 3 --nq 173 --steps 100 --opt qs_resident_kb=6
run qsr_38_nq173 --variant 1 --nq 173 --steps 100
run ts_38_nq16 --variant 2 --nq 16 --steps 100
run ts_38_nq173 --variant 2 --nq 173 --steps 100
echo "=== sustained plain copy"
python - <<'PY'
import torch, time
a=torch.empty(1<<30,dtype=torch.bfloat16,device="cuda"); b=torch.empty_like(a)
for _ in range(3): b.copy_(a)
torch.cuda.synchronize()
for reps in (10, 1000):
    e0=torch.cuda.Event(enable_timing=True); e1=torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps): b.copy_(a)
    e1.record(); torch.cuda.synchronize()
    ms=e0.elapsed_time(e1)/reps
    print("copy reps",reps,"GB/s (read+write)",round(2*a.numel()*2/ms/1e6,1),"ms",round(ms,3))
# read-only sustained: sum reduction
x=torch.empty(1<<31,dtype=torch.bfloat16,device="cuda").zero_()
for reps in (5, 300):
    e0=torch.cuda.Event(enable_timing=True); e1=torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps): s=x.view(torch.int32).sum()
    e1.record(); torch.cuda.synchronize()
    ms=e0.elapsed_time(e1)/reps
    print("read-only sum reps",reps,"GB/s",round(x.numel()*2/ms/1e6,1),"ms",round(ms,3))
PY
echo "=== ncu --set full"
S=4829565
prof() { # out, kernel regex, args...
  out=$1; shift; kr=$1; shift
  timeout 600 ncu --set full --clock-control none --import-source on -k regex:$kr -s 3 -c 1 -o gpurun_out/$out python bench.py --rows $S --steps 1 --warmup 3 --no-cpu-baseline --no-check "$@" > gpurun_out/$out.log 2>&1; echo "$out rc=$?"
}
prof r3b_prof_qs6_4p8 umma_qs --variant 3 --opt qs_resident_kb=6
prof r3b_prof_qs0_4p8 umma_qs --variant 3
prof r3b_prof_qsr_4p8 umma_qs --variant 1
prof r3b_prof_ts_4p8 umma_score --variant 2
prof r3b_prof_qs6_4p8_aniso umma_qs --variant 3 --opt qs_resident_kb=6 --data aniso
prof r3b_prof_fin_aniso finalize --variant 3 --data aniso
ls -la gpurun_out/*.ncu-rep
