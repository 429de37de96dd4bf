"""Resident-load throughput: how fast a collection gets from block files
into device-resident shards, the reference's way and the engine's way, on the same files.

  reference path : `pickle.load` of `passage__emb_p__data_obj_{b}.pb` + `index.add(block)` — what
                   drivers/run_convdr_inference.py:161-180 does for every block of every run
                   (pageable host array -> synchronous H2D).  Timed through convdr_b200.blocks.read_block
                   and FlatIPIndex.add.
  native path    : flat shard files -> `b2f_add_flat_file` (reader threads -> pinned ring -> PCIe), one host
                   thread per GPU (convdr_b200.blocks.load_flat_into).

Usage: python tools/load_bench.py N_GPUS ROWS_PER_SHARD [DIR]   -> one JSON line (files are removed afterwards).
Files are written to DIR (default /dev/shm when it has room, else /tmp) and read back hot from the page cache:
the figure is the memory -> device pipeline, not the disk.
"""
import json
import os
import shutil
import sys
import tempfile
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from convdr_b200 import FlatIPIndex, blocks  # noqa: E402

CAST_GB = 38_636_520 * 3072 / 1e9


def synth_rows(first, n):
    from oracle import c_oracle       # test infrastructure: only used to fabricate input files quickly
    return c_oracle.synth_block(first, n)


def main():
    g = int(sys.argv[1])
    rows = int(sys.argv[2])
    need = g * rows * 3080 * 2.1
    base = sys.argv[3] if len(sys.argv) > 3 else ("/dev/shm" if shutil.disk_usage("/dev/shm").free > need * 1.2 else "/tmp")
    d = tempfile.mkdtemp(prefix="b2f_load_", dir=base)
    out = {"gpus": g, "rows_per_shard": rows, "dir": base, "gbytes": g * rows * 3080 / 1e9, "host_cores": os.cpu_count()}
    try:
        t0 = time.perf_counter()
        for b in range(g):
            P = synth_rows(b * rows, rows)
            ids = blocks.strided_offsets(g * rows, b, g)
            blocks.write_block(d, b, P, ids)
            blocks.write_flat_shard(os.path.join(d, blocks.FLAT_NAME % b), P, ids)
            del P
        out["write_seconds"] = round(time.perf_counter() - t0, 2)

        # ---- reference path: unpickle + add, block after block (the reference resets between blocks; the
        # resident variant keeps them, which does not change the load cost)
        idx = FlatIPIndex(768, devices=list(range(g)))
        idx.reserve((rows + g - 1) // g * g)
        t0 = time.perf_counter()
        t_unpickle = t_add = 0.0
        for b in range(g):
            t1 = time.perf_counter()
            emb, embid = blocks.read_block(d, b)
            t2 = time.perf_counter()
            idx.add(emb)
            t_add += time.perf_counter() - t2
            t_unpickle += t2 - t1
            del emb, embid
        ref_s = time.perf_counter() - t0
        out["reference_path"] = {"seconds": round(ref_s, 3), "unpickle_seconds": round(t_unpickle, 3),
                                 "index_add_seconds": round(t_add, 3), "gb_per_s": out["gbytes"] / ref_s,
                                 "cast_collection_seconds": round(CAST_GB / (out["gbytes"] / ref_s), 1)}
        idx.close()

        # ---- native path
        idx = FlatIPIndex(768, devices=list(range(g)))
        paths = blocks.flat_shard_paths(d, g)
        t0 = time.perf_counter()
        n = blocks.load_flat_into(idx, paths)
        nat_s = time.perf_counter() - t0
        assert n == g * rows and idx.ntotal == n
        out["native_path"] = {"seconds": round(nat_s, 3), "gb_per_s": out["gbytes"] / nat_s,
                              "gb_per_s_per_gpu": out["gbytes"] / nat_s / g,
                              "cast_collection_seconds": round(CAST_GB / (out["gbytes"] / nat_s), 1)}
        # second load into the already sized shards (no allocation, page cache certainly hot)
        idx.reset()
        t0 = time.perf_counter()
        blocks.load_flat_into(idx, paths)
        nat2 = time.perf_counter() - t0
        out["native_path"]["second_load_gb_per_s"] = out["gbytes"] / nat2
        # the labels and rows arrived: spot check
        rows0, ids0 = blocks.open_flat_shard(paths[0])
        got = idx.reconstruct_n(0, 16, shard=0)
        out["rows_match"] = bool(np.array_equal(got, np.asarray(rows0[:16])))
        idx.close()
        out["speedup"] = ref_s / nat_s
    finally:
        shutil.rmtree(d, ignore_errors=True)
    print(json.dumps(out))


if __name__ == "__main__":
    main()
