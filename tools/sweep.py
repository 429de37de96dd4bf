"""Batch-size sweep on one GPU (BASELINE.json config 5 shape): ms per search and streamed GB/s for the
SIMT scan and the tensor engine, device-resident queries, searches queued back to back.  Usage: python tools/sweep.py ROWS K OUT.json"""
import json
import os
import sys
import time

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from convdr_b200 import FlatIPIndex, synth  # noqa: E402


def main():
    rows, k, out = int(sys.argv[1]), int(sys.argv[2]), sys.argv[3]
    nqs = [int(x) for x in sys.argv[4].split(",")] if len(sys.argv) > 4 else [1, 2, 4, 8, 16, 32, 64, 128, 173, 256, 512, 1024]
    dev = torch.device("cuda", 0)
    idx = FlatIPIndex(768)
    idx.set_option("profile", 1)
    idx.reserve(rows)
    for a in range(0, rows, 1 << 22):
        idx.add_synthetic(min(1 << 22, rows - a), first_row=a)
    stream = torch.cuda.ExternalStream(idx.stream_ptr(0), device=dev)
    res = []
    qall = torch.from_numpy(synth.block(0, max(nqs), stream=1)).to(dev)
    for nq in nqs:
        q = qall[:nq].contiguous()
        D = torch.empty((nq, k), dtype=torch.float32, device=dev)
        I = torch.empty((nq, k), dtype=torch.int64, device=dev)
        for path in ("scan_f32", "umma_bf16"):
            if path == "scan_f32" and nq > 64:
                continue
            idx.set_option("path", path)
            reps = 3 if (path == "scan_f32" and nq > 16) else 8
            for _ in range(2):
                idx.search_device_into(q, k, D, I)
            idx.reset_stats()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(stream)
            for _ in range(reps):
                idx.search_device_async(q, k, D, I)     # queued back to back, settled once
            idx.finish()
            e1.record(stream)
            torch.cuda.synchronize()
            sc = idx.stat("score_ms")
            ms = e0.elapsed_time(e1) / reps
            bpr = 3072 if path == "scan_f32" else 1536
            passes = idx.stat("passes") / reps
            r = dict(nq=nq, k=k, rows=rows, path=path, ms=ms, qps=nq / ms * 1e3, score_ms=sc / reps, passes=passes,
                     streamed_gbs=rows * bpr * passes / (sc / reps * 1e-3) / 1e9 if sc > 0 else None,
                     useful_tflops=2.0 * nq * rows * 768 / (ms * 1e-3) / 1e12,
                     fallback=idx.stat("fallback_queries"))
            res.append(r)
            print(json.dumps(r), flush=True)
    json.dump(res, open(out, "w"), indent=1)


if __name__ == "__main__":
    main()
