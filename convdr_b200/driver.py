"""Host-side mirror of the retrieval step of ConvDR's inference driver.

`search_one_by_one` keeps the reference's interface and output contract
(drivers/run_convdr_inference.py:157-242): same arguments, block files, block order, bare stop at
the first missing block, local-index -> offset translation, `>=` merge, `[nq, 2*topN]` float64 /
int64 outputs — so `EvalDevQuery` (:21-113) and everything after it run untouched.  The index object
is the B200 engine (faiss_compat.IndexFlatIP or a multi-GPU clone).

`search_resident` is the "search all at once" variant README.md:216 invites:
blocks are loaded once into device-resident shards with their offsets as labels; one search
returns the global top-k directly (`[nq, topN]`), no per-block add/reset churn and no Python merge.

No arithmetic happens here: scoring, selection and (for search_resident) id translation and merging
are libb2f kernels.  The array bookkeeping of the reference's Python merge is kept in numpy.
"""
from __future__ import annotations

import logging
import time

import numpy as np

logger = logging.getLogger(__name__)

from .blocks import EMB_NAME, EMBID_NAME, flat_shard_paths, load_flat_into, read_block  # noqa: F401


def load_block(ann_data_dir: str, block_id: int):
    """Reader side of the per-rank pickle block format (reference :161-175).  Raises on failure."""
    return read_block(ann_data_dir, block_id)


def _merge_sorted(run_s, run_i, cur_s, cur_i, topN):
    """2-way merge of reference :206-229: first topN of the running list vs the block's topN,
    ties keep the running entry first, both tails drained (output 2*topN wide)."""
    cat_s = np.concatenate([run_s[:, :topN], cur_s[:, :topN]], axis=1)
    cat_i = np.concatenate([run_i[:, :topN], cur_i[:, :topN]], axis=1)
    order = np.argsort(-cat_s, axis=1, kind="stable")
    return np.take_along_axis(cat_s, order, axis=1), np.take_along_axis(cat_i, order, axis=1)


def search_one_by_one(ann_data_dir, gpu_index, query_embedding, topN, max_blocks: int = 8, verbose: bool = True):
    """Drop-in for reference `search_one_by_one(ann_data_dir, gpu_index, query_embedding, topN)`."""
    run_s = run_i = None
    for block_id in range(max_blocks):
        logger.info("Loading passage reps " + str(block_id))
        try:
            passage_embedding, passage_embedding2id = load_block(ann_data_dir, block_id)
        except FileNotFoundError:
            break  # the reference's bare `except: break`; other load errors are surfaced
        gpu_index.add(passage_embedding)
        ts = time.time()
        D, I = gpu_index.search(query_embedding, topN)
        elapsed = time.time() - ts
        if verbose:
            print({"total": elapsed, "data": query_embedding.shape[0],
                   "per_query": elapsed / max(query_embedding.shape[0], 1)})
        cur_i = np.asarray(passage_embedding2id)[I].astype(np.int64)  # -1 wraps to the last offset, as in the reference
        cur_s = D.astype(np.float64)
        gpu_index.reset()
        del passage_embedding, passage_embedding2id
        if run_s is None:
            run_s, run_i = cur_s, cur_i
        else:
            run_s, run_i = _merge_sorted(run_s, run_i, cur_s, cur_i, topN)
    if run_s is None:
        raise TypeError("'NoneType' object is not iterable")  # what the reference raises when no block exists
    return run_s, run_i


def search_resident(ann_data_dir, index, query_embedding, topN, max_blocks: int = 8):
    """Load every block once (labels = the block's offsets), then a single exact search.

    Flat shards (`passage_shard_{b}.b2f`, convdr_b200/blocks.py) are preferred when present: they are
    memory-mapped and streamed in chunks instead of unpickled whole.  Otherwise the reference's pickle
    blocks are read.  Returns (D float64 [nq, topN], I int64 [nq, topN]) — the first topN columns of
    what `search_one_by_one` returns (up to the documented tie order), which is all `EvalDevQuery`
    reads.
    """
    if index.ntotal == 0:
        flat = flat_shard_paths(ann_data_dir, max_blocks)
        if flat:
            load_flat_into(index, flat)
        else:
            import os
            n_files = 0
            while n_files < max_blocks and os.path.exists(os.path.join(ann_data_dir, EMB_NAME % n_files)):
                n_files += 1
            n_blocks = 0
            for block_id in range(max_blocks):
                try:
                    emb, embid = load_block(ann_data_dir, block_id)
                except FileNotFoundError:
                    break
                if n_blocks == 0 and n_files > 1 and hasattr(index, "reserve"):
                    # the blocks of one generation run have (nearly) equal sizes (strided split, utils/util.py:422-424):
                    # size every shard once instead of regrowing it with a device-to-device copy per block
                    shards = getattr(index, "num_shards", 1)
                    index.reserve(-(-(emb.shape[0] + 1) * n_files // shards) + 1)
                index.add_with_ids(emb, np.asarray(embid, dtype=np.int64))
                n_blocks += 1
            if n_blocks == 0:
                raise FileNotFoundError("no passage embedding block under " + str(ann_data_dir))
    D, I = index.search(query_embedding, topN)
    return D.astype(np.float64), I
