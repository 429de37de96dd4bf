#!/bin/bash
# round 2, multi-GPU call: N = number of visible GPUs.  The whole GPU suite (so the two multi-GPU tests run, the
# torchrun one on every GPU), then the headline bench at N with every check (full-size oracle, overflow protocol,
# in-process layout), optionally more configs.  Logs are copied to profiles/ as evidence.
N=${1:-2}
mkdir -p gpurun_out
nvidia-smi -L | head -8
if [ "$3" = "fewtests" ]; then   # N GPU-minutes per minute: only the tests that need more than one GPU
echo "=== multi-GPU tests ($N GPUs visible)"
timeout 600 python -m pytest tests -q -m gpu --timeout 500 -rs -k "peer_memory or multi_gpu" > gpurun_out/r3m_tests_n$N.log 2>&1; echo "rc=$?"; tail -6 gpurun_out/r3m_tests_n$N.log
else
echo "=== gpu tests (all, $N GPUs visible)"
timeout 700 python -m pytest tests -q -m gpu --timeout 500 -rs > gpurun_out/r3m_tests_n$N.log 2>&1; echo "rc=$?"; tail -6 gpurun_out/r3m_tests_n$N.log
fi
bench() { # tag, args...
  tag=$1; shift
  timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N "$@" > gpurun_out/r3m_${tag}_n$N.json 2> gpurun_out/r3m_${tag}_n$N.err; rc=$?
  python - <<PY
import json
try:
    j=json.load(open("gpurun_out/r3m_${tag}_n$N.json"))
    if "sweep" in j:
        for p in j["sweep"]: print("  nq",p["nq"],"ms",round(p["ms_per_search"],3),"q/s",round(p["queries_per_s"]),"hbm",round(p["hbm_frac"],3),"tensor",round(p["tensor_frac"],3),p["bound"],"fb",p["fallback_queries"])
        print("$tag crossover", j["config"]["first_batch_size_where_the_tensor_fraction_exceeds_the_hbm_fraction"])
    else:
        r=j["roofline"]; c=j["clocks"]; s=j.get("sustained") or {}
        print("$tag N=$N rc=$rc ms/step",round(j["ms_per_step"],3),"q/s",round(j["value"]),"e2e",round(j["e2e"]["value"]),round(j["e2e"]["ms_per_step"],3),"kern GB/s",round(r["achieved"]),"frac",round(r["frac"],3),"ms/launch",round(r["ms_per_launch"],3),"sel",round(r["select_kernels_ms_per_step"],3),"clk",c.get("sm_mhz"),c.get("reasons"),"| sus",round(s.get("ms_per_step",0),3),"q/s",round(s.get("value",0)),"| inproc",j.get("inproc"),"| per_rank",j.get("per_rank"),"| check",j.get("check"))
except Exception as e:
    print("$tag rc=$rc FAILED", e); print(open("gpurun_out/r3m_${tag}_n$N.err").read()[-2500:])
PY
}
bench headline --steps 20 --warmup 5
if [ "$2" = "more" ]; then
bench aniso --steps 20 --warmup 5 --data aniso --inproc 0
bench c3 --config c3 --steps 5 --warmup 3 --inproc 0
bench c2 --config c2 --steps 20 --warmup 3 --inproc 0
bench c5 --config c5 --sweep-nq 1,8,64,173,256,1024,4096,16384
echo "=== loader, $N GPUs"
timeout 400 python tools/load_bench.py $N 500000 > gpurun_out/r3m_load_n$N.json 2> gpurun_out/r3m_load_n$N.err; echo "rc=$?"; cat gpurun_out/r3m_load_n$N.json; tail -3 gpurun_out/r3m_load_n$N.err
fi
