"""Batch-size sweep on one GPU: ms per search for AUTO and for every engine forced (TS tensor kernel, QS tensor
kernel, fp32 SIMT scan), device-resident queries, searches queued back to back.  Shows that AUTO is never slower
than the best engine at any batch size.
Usage: python tools/sweep.py ROWS K OUT.json [nq,nq,...]"""
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from convdr_b200 import FlatIPIndex, synth  # noqa: E402

ENGINES = {"auto": dict(path="auto", umma_variant=0), "ts": dict(path="umma_bf16", umma_variant=2),
           "qs": dict(path="umma_bf16", umma_variant=3), "scan_f32": dict(path="scan_f32", umma_variant=0)}


def main():
    rows, k, out = int(sys.argv[1]), int(sys.argv[2]), sys.argv[3]
    nqs = [int(x) for x in sys.argv[4].split(",")] if len(sys.argv) > 4 else [1, 2, 4, 8, 16, 64, 128, 173, 208, 256, 512]
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
        row = {"nq": nq, "k": k, "rows": rows}
        for name, opts in ENGINES.items():
            if name == "scan_f32" and nq > 16:
                continue
            if name == "qs" and nq > 256:
                continue
            for key, val in opts.items():
                idx.set_option(key, val)
            reps = 3 if name == "scan_f32" else 10
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
            row[name + "_ms"] = e0.elapsed_time(e1) / reps
            row[name + "_fallback"] = idx.stat("fallback_queries")
            if name == "auto":
                row["auto_qs_passes"] = idx.stat("qs_passes") / reps
                row["auto_passes"] = idx.stat("passes") / reps
        best = min(v for kk, v in row.items() if kk.endswith("_ms") and kk != "auto_ms")
        row["auto_over_best"] = row["auto_ms"] / best
        res.append(row)
        print(json.dumps(row), flush=True)
    json.dump(res, open(out, "w"), indent=1)


if __name__ == "__main__":
    main()
