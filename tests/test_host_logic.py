"""CPU tests of the host-side mirror of the reference interface: facade names, the shim module,
the driver's merge bookkeeping, shard partitioning, and the one-process-per-GPU plumbing over gloo."""
import os
import socket
import subprocess
import sys
import tempfile

import numpy as np
import pytest

import convdr_b200.faiss_compat as faiss
from convdr_b200 import driver, synth
from convdr_b200.dist import shard_range
from oracle import flat_ip

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_facade_exposes_every_name_the_driver_uses():
    # drivers/run_convdr_inference.py:327-368
    for name in ["get_num_gpus", "StandardGpuResources", "IndexFlatIP", "GpuMultipleClonerOptions",
                 "GpuResourcesVector", "Int32Vector", "index_cpu_to_gpu_multiple"]:
        assert hasattr(faiss, name), name
    res = faiss.StandardGpuResources()
    res.setTempMemory(1 << 20)
    co = faiss.GpuMultipleClonerOptions()
    co.shard = True
    co.usePrecomputed = False
    vres, vdev = faiss.GpuResourcesVector(), faiss.Int32Vector()
    for i in range(3):
        vdev.push_back(i)
        vres.push_back(faiss.StandardGpuResources())
    cpu_index = faiss.IndexFlatIP(768)
    assert cpu_index.d == 768 and cpu_index.ntotal == 0 and cpu_index.is_trained
    idx = faiss.index_cpu_to_gpu_multiple(vres, vdev, cpu_index, co)
    assert idx._devices == [0, 1, 2] and idx.ntotal == 0
    with pytest.raises(RuntimeError):
        faiss.IndexFlatIP(128)
    with pytest.raises(AssertionError):
        cpu_index.add(np.zeros((3, 100), dtype=np.float32))


def test_shim_makes_plain_import_faiss_resolve_to_the_engine():
    code = "import faiss, convdr_b200.faiss_compat as f; assert faiss.IndexFlatIP is f.IndexFlatIP; " \
           "assert faiss.get_num_gpus() >= 0; print('ok')"
    env = dict(os.environ, PYTHONPATH=os.path.join(ROOT, "convdr_b200", "shim"))
    out = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, env=env, cwd="/tmp")
    assert out.returncode == 0 and "ok" in out.stdout, out.stderr


def golden(name):
    z = np.load(os.path.join(ROOT, "tests", "golden", "search_one_by_one.npz"))
    return z[name + "/D"], z[name + "/I"], z[name + "/params"]


@pytest.mark.parametrize("name", ["three_blocks", "one_block", "eight_blocks_k100",
                                  "short_block_wraps_minus_one", "ties_across_blocks"])
def test_driver_mirror_reproduces_reference_golden_with_a_cpu_index(name):
    """driver.search_one_by_one's bookkeeping (block loop, offset translation incl. -1 wrap, `>=`
    merge, 2*topN width, float64) against vectors from the reference's own function.  The index
    handed in is the oracle's CPU IndexFlatIP — the test isolates the host logic."""
    Dg, Ig, params = golden(name)
    n, W, nq, topN, dup = (int(v) for v in params)
    P = synth.block(0, n, seed=7, stream=0)
    if dup:
        P[n - dup:] = P[:dup]
    Q = synth.block(0, nq, seed=7, stream=1)
    with tempfile.TemporaryDirectory() as tmp:
        for r in range(W):
            ids = np.arange(r, n, W, dtype=np.int64)
            flat_ip.write_block(tmp, r, P[ids], ids)
        D, I = driver.search_one_by_one(tmp, flat_ip.IndexFlatIP(768), Q, topN, verbose=False)
    assert D.dtype == np.float64 and I.dtype == np.int64 and D.shape == Dg.shape
    np.testing.assert_array_equal(I, Ig)
    np.testing.assert_array_equal(D, Dg)


def test_driver_mirror_surfaces_corrupt_blocks_instead_of_swallowing_them():
    with tempfile.TemporaryDirectory() as tmp:
        flat_ip.write_block(tmp, 0, synth.block(0, 10), np.arange(10))
        with open(os.path.join(tmp, driver.EMB_NAME % 1), "wb") as f:
            f.write(b"not a pickle")
        with open(os.path.join(tmp, driver.EMBID_NAME % 1), "wb") as f:
            f.write(b"not a pickle")
        with pytest.raises(Exception):
            driver.search_one_by_one(tmp, flat_ip.IndexFlatIP(768), synth.block(0, 2, stream=1), 3, verbose=False)
    with tempfile.TemporaryDirectory() as tmp, pytest.raises(TypeError):
        driver.search_one_by_one(tmp, flat_ip.IndexFlatIP(768), synth.block(0, 2, stream=1), 3, verbose=False)


def test_shard_range_is_a_contiguous_partition():
    for n in [0, 1, 7, 100, 38636520]:
        for w in [1, 2, 4, 8]:
            edges = [shard_range(n, r, w) for r in range(w)]
            assert edges[0][0] == 0 and edges[-1][1] == n
            assert all(edges[i][1] == edges[i + 1][0] for i in range(w - 1))
            sizes = [b - a for a, b in edges]
            assert max(sizes) - min(sizes) <= 1


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


@pytest.mark.parametrize("world", [2])
def test_sharded_search_over_gloo_world2_equals_single_index(world):
    """N>1 path on CPU: row partition + all-gather + merge, world_size 2 over gloo.  The rank-local
    search and the merge are oracle stand-ins here (the CUDA ones are covered by -m gpu)."""
    port = _free_port()
    script = os.path.join(ROOT, "tests", "_gloo_worker.py")
    procs = []
    with tempfile.TemporaryDirectory() as tmp:
        for r in range(world):
            env = dict(os.environ, RANK=str(r), WORLD_SIZE=str(world), MASTER_ADDR="127.0.0.1",
                       MASTER_PORT=str(port), OUT_DIR=tmp, PYTHONPATH=ROOT)
            procs.append(subprocess.Popen([sys.executable, script], env=env, stdout=subprocess.PIPE,
                                          stderr=subprocess.STDOUT, text=True))
        outs = [p.communicate(timeout=180)[0] for p in procs]
        assert all(p.returncode == 0 for p in procs), "\n".join(outs)
        P = synth.block(0, 5003, seed=11)
        Q = synth.block(0, 9, seed=11, stream=1)
        D_ref, I_ref = flat_ip.knn_inner_product(Q, P, 25)
        for r in range(world):
            z = np.load(os.path.join(tmp, f"rank{r}.npz"))
            np.testing.assert_array_equal(z["I"], I_ref)
            np.testing.assert_allclose(z["D"], D_ref, rtol=1e-6)


def test_weighted_shard_ranges_partition_exactly_and_follow_the_speeds():
    from convdr_b200.dist import balance_weights
    times = [1.00, 1.08, 0.95, 1.02, 1.40, 1.00, 0.97, 1.01]       # one straggler far outside the clamp
    w = balance_weights(times)
    assert min(w) >= 0.85 - 1e-12 and max(w) <= 1.15 + 1e-12
    n = 38_636_520
    cuts = [shard_range(n, r, 8, w) for r in range(8)]
    assert cuts[0][0] == 0 and cuts[-1][1] == n
    assert all(cuts[r][1] == cuts[r + 1][0] for r in range(7))
    sizes = [b - a for a, b in cuts]
    assert sizes[2] > sizes[0] > sizes[1] > sizes[4]                  # faster GPUs hold more rows
    assert [shard_range(1001, r, 4, None) for r in range(4)] == [shard_range(1001, r, 4) for r in range(4)]


def test_bench_reference_arm_prints_one_contract_json_line():
    """`bench.py --impl reference` (the CPU arm the driver runs next to ours) on a tiny sample: exactly one
    JSON line on stdout with the keys of the bench contract."""
    import json
    res = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--cpu-sample-rows",
                          "20000", "--steps", "1", "--warmup", "1"], capture_output=True, text=True, timeout=300)
    assert res.returncode == 0, res.stderr[-2000:]
    lines = [ln for ln in res.stdout.splitlines() if ln.strip()]
    assert len(lines) == 1
    j = json.loads(lines[0])
    assert j["impl"] == "reference" and j["unit"] == "queries/s" and j["higher_is_better"] is True
    assert j["value"] > 0 and j["e2e"]["value"] == j["value"] and j["e2e"]["h2d_bytes_per_step"] == 0
    assert j["cpu_baseline"]["kind"] == "port" and j["cpu_baseline"]["cores"] >= 1 and "sample" in j["cpu_baseline"]
    assert j["metric"].startswith("exact top-100 queries/sec") and j["config"]["workload"]


def test_bench_configs_follow_baseline_json():
    """bench.py --config cN sets the shape BASELINE.json's configs[N-1] names; c4 is the headline metric."""
    import importlib
    sys.argv, saved = ["bench.py", "--config", "c3"], sys.argv
    try:
        bench = importlib.import_module("bench")
        args = bench.parse_args()
        assert (args.rows, args.nq, args.k) == (11_100_000, 5571, 100)
        assert "11.1M" in bench.metric_name(args)
        sys.argv = ["bench.py"]
        args = bench.parse_args()
        assert (args.rows, args.nq, args.k) == (38_636_520, 173, 100) and bench.metric_name(args) == bench.METRIC
        sys.argv = ["bench.py", "--config", "c2"]
        args = bench.parse_args()
        assert (args.rows, args.nq, args.k) == (8_841_823, 173, 1000)
    finally:
        sys.argv = saved


def test_flat_shard_rows_and_fallback_loader(tmp_path):
    from convdr_b200 import blocks
    P = np.arange(5 * 768, dtype=np.float32).reshape(5, 768)
    path = blocks.write_flat_shard(str(tmp_path / (blocks.FLAT_NAME % 0)), P, np.arange(5, dtype=np.int64) * 2)
    assert blocks.flat_shard_rows(path) == 5

    class Fake:                       # no add_flat_file: the chunked add_with_ids path
        def __init__(self):
            self.rows, self.ids = [], []
        def add_with_ids(self, x, ids):
            self.rows.append(x.copy()); self.ids.append(ids.copy())
    f = Fake()
    assert blocks.load_flat_into(f, [path], chunk_rows=2) == 5
    np.testing.assert_array_equal(np.concatenate(f.rows), P)
    np.testing.assert_array_equal(np.concatenate(f.ids), np.arange(5) * 2)
