"""Passage-embedding block files: the reference's per-rank pickle format and a flat shard format.

Reference side:
  * writer — `barrier_array_merge` dumps each rank's arrays with `pickle.dump(..., protocol=4)` to
    `{prefix}_data_obj_{rank}.pb` (utils/util.py:105-111); `StreamInferenceDoc` calls it with the
    prefixes `passage__emb_p_` / `passage__embid_p_` (drivers/gen_passage_embeddings.py:146-169);
  * reader — `search_one_by_one` unpickles `passage__emb_p__data_obj_{b}.pb` (float32 [n, 768]) and
    `passage__embid_p__data_obj_{b}.pb` (int64 [n], global passage offsets) for b = 0..7 and stops
    at the first missing file (drivers/run_convdr_inference.py:161-177).

`write_block` / `read_block` keep that format byte-compatible, so blocks written here are read by
the unmodified reference and vice versa.

The flat format is what README.md:216 ("search all at once") needs in practice: unpickling 118 GB
per run is the reference's real wall-clock cost.  A flat shard is ONE file

    header  (64 bytes)  magic "B2FSHARD", version, d, n, dtype code, offsets
    rows    float32 [n, d]   row-major, 64-byte aligned
    ids     int64   [n]      global passage offsets, 64-byte aligned

that the engine streams into a shard through pinned staging buffers (`load_flat_into` ->
`b2f_add_flat_file`: reader threads, two copies in flight per GPU, all GPUs at once), so host memory
stays bounded and the H2D copies run at PCIe speed.  No arithmetic happens here; the bf16 shadow and
the norm bounds are built on the GPU by the ingest kernel (~1 ms per million rows: putting the shadow
in the file would add 50 % more PCIe bytes to save that millisecond, so it stays out).
"""
from __future__ import annotations

import os
import pickle
import struct
from typing import Iterator, Optional, Tuple

import numpy as np

EMB_NAME = "passage__emb_p__data_obj_%d.pb"      # utils/util.py:108-109 + gen_passage_embeddings.py:158
EMBID_NAME = "passage__embid_p__data_obj_%d.pb"  # gen_passage_embeddings.py:164
FLAT_NAME = "passage_shard_%d.b2f"

MAGIC = b"B2FSHARD"
VERSION = 1
HEADER_BYTES = 64
_HEADER = struct.Struct("<8sIIQIQQ")   # magic, version, d, n, dtype code (0 = float32), rows offset, ids offset
_ALIGN = 64


# ------------------------------------------------------------------------------------------------
# the reference's pickle blocks
# ------------------------------------------------------------------------------------------------
def write_block(output_dir: str, rank: int, embedding: np.ndarray, embedding2id: np.ndarray) -> Tuple[str, str]:
    """What rank `rank` of gen_passage_embeddings.py leaves on disk (pickle protocol 4 of a float32
    [n, 768] array and of an int64 [n] array).  Returns the two paths."""
    embedding = np.ascontiguousarray(embedding, dtype=np.float32)
    embedding2id = np.ascontiguousarray(embedding2id, dtype=np.int64)
    assert embedding.ndim == 2 and embedding2id.shape == (embedding.shape[0],)
    os.makedirs(output_dir, exist_ok=True)
    p_emb = os.path.join(output_dir, EMB_NAME % rank)
    p_id = os.path.join(output_dir, EMBID_NAME % rank)
    with open(p_emb, "wb") as handle:
        pickle.dump(embedding, handle, protocol=4)
    with open(p_id, "wb") as handle:
        pickle.dump(embedding2id, handle, protocol=4)
    return p_emb, p_id


def read_block(ann_data_dir: str, block_id: int) -> Tuple[np.ndarray, np.ndarray]:
    """Reader of reference :161-175.  Raises FileNotFoundError for a missing block."""
    with open(os.path.join(ann_data_dir, EMB_NAME % block_id), "rb") as handle:
        emb = pickle.load(handle)
    with open(os.path.join(ann_data_dir, EMBID_NAME % block_id), "rb") as handle:
        embid = pickle.load(handle)
    return emb, embid


def iter_blocks(ann_data_dir: str, max_blocks: int = 8) -> Iterator[Tuple[int, np.ndarray, np.ndarray]]:
    """Blocks 0, 1, ... until the first missing one (the reference's `except: break`, :176-177)."""
    for b in range(max_blocks):
        try:
            emb, embid = read_block(ann_data_dir, b)
        except FileNotFoundError:
            return
        yield b, emb, embid


def strided_offsets(n_total: int, rank: int, world: int) -> np.ndarray:
    """Global offsets held by rank `rank` when `world` ranks generated the collection: the
    StreamingDataset stride b, b+W, b+2W, ... (`i % num_replicas != rank`, utils/util.py:422-424)."""
    return np.arange(rank, n_total, world, dtype=np.int64)


# ------------------------------------------------------------------------------------------------
# flat shards
# ------------------------------------------------------------------------------------------------
def _round_up(v: int, m: int) -> int:
    return (v + m - 1) // m * m


def write_flat_shard(path: str, embedding: np.ndarray, embedding2id: np.ndarray) -> str:
    """Write one flat shard file (layout in the module docstring)."""
    embedding = np.ascontiguousarray(embedding, dtype=np.float32)
    embedding2id = np.ascontiguousarray(embedding2id, dtype=np.int64)
    assert embedding.ndim == 2 and embedding2id.shape == (embedding.shape[0],)
    n, d = embedding.shape
    rows_off = HEADER_BYTES
    ids_off = _round_up(rows_off + n * d * 4, _ALIGN)
    tmp = path + ".tmp"
    with open(tmp, "wb") as f:
        f.write(_HEADER.pack(MAGIC, VERSION, d, n, 0, rows_off, ids_off).ljust(HEADER_BYTES, b"\0"))
        f.write(embedding.tobytes(order="C"))
        f.write(b"\0" * (ids_off - (rows_off + n * d * 4)))
        f.write(embedding2id.tobytes(order="C"))
    os.replace(tmp, path)
    return path


def open_flat_shard(path: str) -> Tuple[np.ndarray, np.ndarray]:
    """Memory-map a flat shard: (rows float32 [n, d], ids int64 [n]), both read-only views."""
    with open(path, "rb") as f:
        raw = f.read(HEADER_BYTES)
    if len(raw) < HEADER_BYTES:
        raise ValueError(f"{path}: truncated header")
    magic, version, d, n, dtype_code, rows_off, ids_off = _HEADER.unpack(raw[:_HEADER.size])
    if magic != MAGIC:
        raise ValueError(f"{path}: not a b2f flat shard (bad magic)")
    if version != VERSION or dtype_code != 0:
        raise ValueError(f"{path}: unsupported version {version} / dtype code {dtype_code}")
    size = os.path.getsize(path)
    if ids_off + n * 8 > size or rows_off + n * d * 4 > ids_off:
        raise ValueError(f"{path}: file shorter than its header claims")
    if n == 0:
        return np.zeros((0, d), dtype=np.float32), np.zeros((0,), dtype=np.int64)
    rows = np.memmap(path, dtype=np.float32, mode="r", offset=rows_off, shape=(n, d))
    ids = np.memmap(path, dtype=np.int64, mode="r", offset=ids_off, shape=(n,))
    return rows, ids


def convert_blocks_to_flat(ann_data_dir: str, out_dir: Optional[str] = None, max_blocks: int = 8) -> list:
    """One-time conversion of the reference's pickle blocks into flat shards (one per block)."""
    out_dir = out_dir or ann_data_dir
    os.makedirs(out_dir, exist_ok=True)
    paths = []
    for b, emb, embid in iter_blocks(ann_data_dir, max_blocks):
        paths.append(write_flat_shard(os.path.join(out_dir, FLAT_NAME % b), emb, np.asarray(embid, dtype=np.int64)))
    return paths


def flat_shard_paths(shard_dir: str, max_blocks: int = 8) -> list:
    paths = []
    for b in range(max_blocks):
        p = os.path.join(shard_dir, FLAT_NAME % b)
        if not os.path.exists(p):
            break
        paths.append(p)
    return paths


def load_flat_into(index, paths, chunk_rows: int = 1 << 18, rank: int = 0, world: int = 1) -> int:
    """Stream flat shards into `index` (anything with `add_with_ids(x, ids)`), `chunk_rows` rows at a
    time, labels = the stored passage offsets.  With world > 1 (one process per GPU) shard file i goes
    to rank i % world — the natural mapping of the reference's 8 blocks onto G GPUs.
    Returns the number of rows added by this rank."""
    mine = [p for i, p in enumerate(paths) if i % world == rank]
    if hasattr(index, "add_flat_file"):
        # the engine's native loader: reader threads -> pinned staging -> PCIe, one host thread per shard so
        # the links of all GPUs run at once; shard sizes come from the headers, so every shard is sized once
        from concurrent.futures import ThreadPoolExecutor
        n_shards = index.num_shards
        per_shard = [[] for _ in range(n_shards)]
        rows_per_shard = [0] * n_shards
        for i, p in enumerate(mine):
            per_shard[i % n_shards].append(p)
            rows_per_shard[i % n_shards] += flat_shard_rows(p)
        if index.ntotal == 0 and max(rows_per_shard) > 0:
            index.reserve(max(rows_per_shard))

        def load(s):
            for p in per_shard[s]:
                index.add_flat_file(p, shard=s)
        with ThreadPoolExecutor(max_workers=n_shards) as ex:
            list(ex.map(load, range(n_shards)))
        return sum(rows_per_shard)
    added = 0
    for p in mine:
        rows, ids = open_flat_shard(p)
        for a in range(0, rows.shape[0], chunk_rows):
            b = min(rows.shape[0], a + chunk_rows)
            index.add_with_ids(np.ascontiguousarray(rows[a:b]), np.ascontiguousarray(ids[a:b]))
            added += b - a
    return added


def save_index_to_flat(index, out_dir: str) -> list:
    """One flat shard file per device-resident shard of `index` (`passage_shard_{s}.b2f`), written straight from
    device memory; `load_flat_into` restores the same shards.  The writer side of the format for collections
    that were added from device tensors or synthetic generators."""
    os.makedirs(out_dir, exist_ok=True)
    paths = []
    for s in range(index.num_shards):
        p = os.path.join(out_dir, FLAT_NAME % s)
        index.write_flat_file(p, shard=s)
        paths.append(p)
    return paths


def flat_shard_rows(path: str) -> int:
    """Row count from a flat shard's header."""
    with open(path, "rb") as f:
        raw = f.read(HEADER_BYTES)
    magic, version, d, n, dtype_code, rows_off, ids_off = _HEADER.unpack(raw[:_HEADER.size])
    if magic != MAGIC:
        raise ValueError(f"{path}: not a b2f flat shard (bad magic)")
    return int(n)
