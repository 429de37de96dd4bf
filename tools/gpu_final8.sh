#!/bin/bash
# final-code refresh on 8 GPUs: the multi-GPU tests + merge tests, headline, config 2, a short config-5 sweep
N=8
mkdir -p gpurun_out
timeout 300 python -m pytest tests -q -m gpu --timeout 250 -rs -k "peer_memory or multi_gpu or merge" > gpurun_out/r3m_tests_n$N.log 2>&1; echo "rc=$?"; tail -3 gpurun_out/r3m_tests_n$N.log
bench() { tag=$1; shift
  timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N "$@" > gpurun_out/r3m_${tag}_n$N.json 2> gpurun_out/r3m_${tag}_n$N.err; rc=$?
  python - <<PY
import json
try:
    j=json.load(open("gpurun_out/r3m_${tag}_n$N.json"))
    if "sweep" in j:
        for p in j["sweep"]: print("  nq",p["nq"],"ms",round(p["ms_per_search"],3),"q/s",round(p["queries_per_s"]),"hbm",round(p["hbm_frac"],3),"tensor",round(p["tensor_frac"],3),p["bound"],"fb",p["fallback_queries"])
    else:
        r=j["roofline"]; s=j.get("sustained") or {}
        print("$tag N=$N rc=$rc ms/step",round(j["ms_per_step"],3),"q/s",round(j["value"]),"e2e",round(j["e2e"]["value"]),round(j["e2e"]["ms_per_step"],3),"frac",round(r["frac"],3),"ms/launch",round(r["ms_per_launch"],3),"sel",round(r["select_kernels_ms_per_step"],3),"| sus",round(s.get("ms_per_step",0),3),round(s.get("value",0)),"| inproc",(j.get("inproc") or {}).get("value"),(j.get("inproc") or {}).get("ms_per_step"),"| xchg",j["per_rank"]["exchange_wait_merge_ms_per_step"],"| check",{k:v for k,v in j["check"].items() if k in ("oracle_violations","overflow_redo_ok","fallback_queries","host_api_equals_device_api")})
except Exception as e:
    print("$tag rc=$rc FAILED", e); print(open("gpurun_out/r3m_${tag}_n$N.err").read()[-2500:])
PY
}
bench headline --steps 20 --warmup 5
bench c2 --config c2 --steps 20 --warmup 3 --inproc 0
bench c5 --config c5 --sweep-nq 1,173,1024,16384
