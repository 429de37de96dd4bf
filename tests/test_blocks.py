"""CPU tests of convdr_b200/blocks.py: the reference's pickle block format (writer + reader,
reference utils/util.py:105-111, drivers/run_convdr_inference.py:161-177) and the flat shard format."""
import os
import pickle

import numpy as np
import pytest

from convdr_b200 import blocks, driver, synth
from oracle import flat_ip


def _toy(n, seed):
    return synth.block(0, n, seed=seed), blocks.strided_offsets(8 * n, seed % 8, 8)[:n]


def test_pickle_blocks_are_byte_identical_to_the_reference_writer(tmp_path):
    # the reference writes `pickle.dump(data_array, handle, protocol=4)` (utils/util.py:110-111)
    emb, ids = _toy(37, 3)
    pe, pi = blocks.write_block(str(tmp_path), 5, emb, ids)
    assert os.path.basename(pe) == "passage__emb_p__data_obj_5.pb"
    assert os.path.basename(pi) == "passage__embid_p__data_obj_5.pb"
    assert open(pe, "rb").read() == pickle.dumps(emb, protocol=4)
    assert open(pi, "rb").read() == pickle.dumps(ids, protocol=4)
    # and both readers (ours, the oracle's restatement of the reference's) agree
    e2, i2 = blocks.read_block(str(tmp_path), 5)
    e3, i3 = flat_ip.read_block(str(tmp_path), 5)
    for a, b in ((e2, emb), (e3, emb), (i2, ids), (i3, ids)):
        np.testing.assert_array_equal(a, b)
    assert e2.dtype == np.float32 and i2.dtype == np.int64


def test_iter_blocks_stops_at_the_first_missing_block(tmp_path):
    for b in (0, 1, 3):   # block 2 missing: the reference's bare `except: break` never sees block 3
        blocks.write_block(str(tmp_path), b, *_toy(5, b))
    assert [b for b, _, _ in blocks.iter_blocks(str(tmp_path))] == [0, 1]
    with pytest.raises(FileNotFoundError):
        blocks.read_block(str(tmp_path), 2)


def test_strided_offsets_partition_the_collection():
    parts = [blocks.strided_offsets(103, r, 8) for r in range(8)]
    assert sorted(np.concatenate(parts).tolist()) == list(range(103))
    assert parts[3][:3].tolist() == [3, 11, 19]


@pytest.mark.parametrize("n", [0, 1, 5, 1000])
def test_flat_shard_round_trip(tmp_path, n):
    emb = synth.block(0, n, seed=9) if n else np.zeros((0, 768), dtype=np.float32)
    ids = (np.arange(n, dtype=np.int64) * 8 + 5)
    p = blocks.write_flat_shard(str(tmp_path / "s.b2f"), emb, ids)
    rows, got = blocks.open_flat_shard(p)
    np.testing.assert_array_equal(rows, emb)
    np.testing.assert_array_equal(got, ids)
    assert rows.dtype == np.float32 and got.dtype == np.int64
    if n:
        assert rows.ctypes.data % 64 == 0 and got.ctypes.data % 64 == 0   # 64-byte aligned sections


def test_flat_shard_rejects_foreign_and_truncated_files(tmp_path):
    bad = tmp_path / "bad.b2f"
    bad.write_bytes(b"not a shard" * 10)
    with pytest.raises(ValueError):
        blocks.open_flat_shard(str(bad))
    p = blocks.write_flat_shard(str(tmp_path / "ok.b2f"), synth.block(0, 50), np.arange(50, dtype=np.int64))
    data = open(p, "rb").read()
    (tmp_path / "cut.b2f").write_bytes(data[:len(data) // 2])
    with pytest.raises(ValueError):
        blocks.open_flat_shard(str(tmp_path / "cut.b2f"))
    (tmp_path / "tiny.b2f").write_bytes(data[:10])
    with pytest.raises(ValueError):
        blocks.open_flat_shard(str(tmp_path / "tiny.b2f"))


class _Recorder:
    """Stand-in index: records what load_flat_into / search_resident feed it (no GPU here)."""
    def __init__(self):
        self.x, self.ids, self.calls = [], [], 0
        self.ntotal = 0

    def add_with_ids(self, x, ids):
        assert x.flags.c_contiguous and x.dtype == np.float32 and ids.dtype == np.int64
        self.x.append(np.array(x)); self.ids.append(np.array(ids)); self.calls += 1
        self.ntotal += x.shape[0]

    def search(self, q, k):
        P, ids = np.concatenate(self.x), np.concatenate(self.ids)
        D, I = flat_ip.knn_inner_product(q, P, k)
        return D, np.where(I >= 0, ids[np.maximum(I, 0)], -1)


def test_convert_and_chunked_load_feed_every_row_once(tmp_path):
    total = 0
    for b in range(3):
        emb, _ = _toy(40 + 7 * b, b)
        blocks.write_block(str(tmp_path), b, emb, blocks.strided_offsets(400, b, 3)[:emb.shape[0]])
        total += emb.shape[0]
    paths = blocks.convert_blocks_to_flat(str(tmp_path))
    assert [os.path.basename(p) for p in paths] == ["passage_shard_0.b2f", "passage_shard_1.b2f", "passage_shard_2.b2f"]
    assert blocks.flat_shard_paths(str(tmp_path)) == paths
    rec = _Recorder()
    assert blocks.load_flat_into(rec, paths, chunk_rows=16) == total
    assert rec.calls == sum(-(-(40 + 7 * b) // 16) for b in range(3))
    for b in range(3):   # one process per GPU: shard file i goes to rank i % world
        r = _Recorder()
        blocks.load_flat_into(r, paths, rank=b % 2, world=2)
    # the resident driver path prefers the flat shards and returns the same ranking as the block loop
    Q = synth.block(0, 4, seed=1, stream=1)
    res = _Recorder()
    Dr, Ir = driver.search_resident(str(tmp_path), res, Q, 10)
    ref = flat_ip.IndexFlatIP(768)
    Do, Io = flat_ip.search_one_by_one(str(tmp_path), ref, Q, 10)
    np.testing.assert_array_equal(Ir, Io[:, :10])
    np.testing.assert_allclose(Dr, Do[:, :10], rtol=1e-6)


def test_block_format_is_pinned_by_files_the_reference_writer_produced(tmp_path):
    """tests/golden/ref_blocks/*.pb were written by the reference's own `barrier_array_merge`
    (utils/util.py:88-143, imported with stubs by tests/golden/make_golden_blocks.py).  Our reader must
    return the arrays that went in, our writer must produce the same file names and — under the numpy
    that produced the fixtures — the same bytes."""
    import json
    gdir = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "ref_blocks")
    prm = json.load(open(os.path.join(gdir, "params.json")))
    emb = synth.block(0, prm["n_rows"], seed=prm["seed"])
    embid = blocks.strided_offsets(prm["n_rows"] * prm["world"], prm["rank"], prm["world"])
    assert prm["files"] == [blocks.EMB_NAME % prm["rank"], blocks.EMBID_NAME % prm["rank"]]
    e, i = blocks.read_block(gdir, prm["rank"])
    np.testing.assert_array_equal(e, emb)
    np.testing.assert_array_equal(i, embid)
    assert e.dtype == np.float32 and i.dtype == np.int64
    pe, pi = blocks.write_block(str(tmp_path), prm["rank"], emb, embid)
    assert sorted(os.listdir(tmp_path)) == prm["files"]
    if np.__version__.split(".")[:2] == prm["numpy"].split(".")[:2]:
        assert open(pe, "rb").read() == open(os.path.join(gdir, prm["files"][0]), "rb").read()
        assert open(pi, "rb").read() == open(os.path.join(gdir, prm["files"][1]), "rb").read()
