#!/bin/bash
# merge kernels (valid-prefix searches, 1024 threads for long lists), cached segment upload, bare e2e: tests on 2 GPUs,
# merge micro-timing, N=2 at the 8-GPU shard size, k=1000
mkdir -p gpurun_out
timeout 600 python -m pytest tests -q -m gpu --timeout 400 > gpurun_out/r3j_tests.log 2>&1; echo "rc=$?"; tail -4 gpurun_out/r3j_tests.log
python - <<'PY'
import torch, numpy as np, sys
sys.path.insert(0, ".")
from convdr_b200 import FlatIPIndex
idx = FlatIPIndex(768); idx.add(np.zeros((64, 768), np.float32))
st = torch.cuda.ExternalStream(idx.stream_ptr(0))
for G, k in ((8, 100), (8, 1000), (2, 100)):
    nq = 173
    D = -torch.sort(-torch.rand((G, nq, k), device="cuda"), dim=2).values.contiguous()
    I = torch.randint(0, 1 << 40, (G, nq, k), device="cuda")
    i_off = (nq * k * 4 + 15) // 16 * 16; part = i_off + nq * k * 8
    recv = torch.zeros((G, part), dtype=torch.uint8, device="cuda")
    for g in range(G):
        recv[g, :nq * k * 4] = D[g].reshape(-1).view(torch.uint8)
        recv[g, i_off:] = I[g].reshape(-1).view(torch.uint8)
    Do = torch.empty((nq, k), device="cuda"); Io = torch.empty((nq, k), dtype=torch.int64, device="cuda")
    torch.cuda.synchronize()
    for _ in range(3): idx.merge_packed_device_async(recv, G, part, i_off, nq, k, Do, Io)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(st)
    for _ in range(50): idx.merge_packed_device_async(recv, G, part, i_off, nq, k, Do, Io)
    e1.record(st); torch.cuda.synchronize()
    # reference result
    cat = D.permute(1, 0, 2).reshape(nq, G * k); catI = I.permute(1, 0, 2).reshape(nq, G * k)
    o = torch.argsort(-cat, dim=1, stable=True)[:, :k]
    ok = bool(torch.equal(torch.gather(cat, 1, o), Do) and torch.equal(torch.gather(catI, 1, o), Io))
    print("merge G", G, "k", k, "us per launch", round(e0.elapsed_time(e1) / 50 * 1000, 1), "correct", ok)
PY
b2() { tag=$1; shift
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 20 --warmup 5 --inproc 0 --no-oracle-check "$@" > gpurun_out/r3j_$tag.json 2> gpurun_out/r3j_$tag.err; echo "rc=$?"
python - <<PY
import json
try:
    j=json.load(open("gpurun_out/r3j_$tag.json")); r=j["roofline"]
    print("$tag ms/step",round(j["ms_per_step"],3),"q/s",round(j["value"]),"e2e",round(j["e2e"]["value"]),round(j["e2e"]["ms_per_step"],3),"frac",round(r["frac"],3),"ms/launch",round(r["ms_per_launch"],3),"sel",round(r["select_kernels_ms_per_step"],3),"sus",round(j["sustained"]["ms_per_step"],3),"per_rank",j["per_rank"],"fb",j["check"]["fallback_queries"])
except Exception as e:
    print("$tag FAILED",e); print(open("gpurun_out/r3j_$tag.err").read()[-2000:])
PY
}
b2 n2_4p8each --rows 9659130
b2 n2_k1000 --rows 9659130 --k 1000
