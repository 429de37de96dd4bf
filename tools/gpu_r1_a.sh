#!/bin/bash
# First GPU session of round 1: parity per engine (separate processes so a trap in one engine does
# not poison the others), a sanitizer pass on a tiny case, and two small bench runs.
mkdir -p gpurun_out
nvidia-smi -L
nvidia-smi --query-gpu=memory.total,memory.used --format=csv
python -c "import torch; print(torch.__version__, torch.cuda.is_available())"
free -g | head -2; nproc
echo "=== smoke (per engine)"
for p in scan_f32 scan_exact umma_bf16; do
timeout 300 python - <<PY > gpurun_out/smoke_$p.log 2>&1
import numpy as np, time
from convdr_b200 import FlatIPIndex
from oracle import c_oracle, flat_ip
P = c_oracle.synth_block(0, 20000); Q = c_oracle.synth_block(0, 40, stream=1)
Dt, It = flat_ip.truth_fp64(Q, P, 10)
idx = FlatIPIndex(768); idx.set_option("path", "$p"); idx.add(P)
D, I = idx.search(Q, 10)
so = lambda qi, ids: Q[qi].astype(np.float64) @ P[ids].astype(np.float64).T
print("$p", flat_ip.compare(D, I, Dt, It, so), "launches", idx.stat("launches"), "fallback", idx.stat("fallback_queries"))
print(I[0], It[0]); print(D[0], Dt[0])
PY
echo "smoke $p rc=$?"; tail -5 gpurun_out/smoke_$p.log
done
echo "=== pytest scan engines"
timeout 1200 python -m pytest tests/test_gpu_parity.py -q -m gpu -k "scan_f32 or scan_exact" --maxfail=10 > gpurun_out/t_scan.log 2>&1; echo "scan rc=$?"; tail -15 gpurun_out/t_scan.log
echo "=== pytest umma engine"
timeout 900 python -m pytest tests/test_gpu_parity.py -q -m gpu -k "umma_bf16" --maxfail=6 > gpurun_out/t_umma.log 2>&1; echo "umma rc=$?"; tail -15 gpurun_out/t_umma.log
echo "=== pytest misc"
timeout 900 python -m pytest tests/test_gpu_parity.py -q -m gpu -k "not scan_f32 and not scan_exact and not umma_bf16" --maxfail=10 > gpurun_out/t_misc.log 2>&1; echo "misc rc=$?"; tail -15 gpurun_out/t_misc.log
echo "=== sanitizer (tiny)"
timeout 600 compute-sanitizer --tool memcheck --print-limit 20 python - > gpurun_out/sanitizer.log 2>&1 <<PY
import numpy as np
from convdr_b200 import FlatIPIndex
from oracle import c_oracle
P = c_oracle.synth_block(0, 3000); Q = c_oracle.synth_block(0, 20, stream=1)
for p in ["scan_f32", "scan_exact", "umma_bf16"]:
    idx = FlatIPIndex(768); idx.set_option("path", p); idx.add(P); D, I = idx.search(Q, 10); print(p, I[0][:5])
PY
echo "sanitizer rc=$?"; tail -8 gpurun_out/sanitizer.log
echo "=== bench small"
timeout 600 python bench.py --rows 4000000 --steps 5 --path scan_f32 --nq 8 --no-cpu-baseline > gpurun_out/bench_scan_small.json 2> gpurun_out/bench_scan_small.err; echo "rc=$?"; cat gpurun_out/bench_scan_small.json; tail -3 gpurun_out/bench_scan_small.err
timeout 600 python bench.py --rows 4000000 --steps 5 --no-cpu-baseline > gpurun_out/bench_umma_small.json 2> gpurun_out/bench_umma_small.err; echo "rc=$?"; cat gpurun_out/bench_umma_small.json; tail -3 gpurun_out/bench_umma_small.err
