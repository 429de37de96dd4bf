"""ORACLE — TEST INFRASTRUCTURE ONLY.  Never imported by the product path (convdr_b200/); only
tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may use it.

CPU restatement of the dense-retrieval hot path of thunlp/ConvDR:

  * `IndexFlatIP`              — what `faiss.IndexFlatIP(768)` does for add / search / reset
                                 (reference drivers/run_convdr_inference.py:353, :180, :182, :202)
  * `search_one_by_one`        — the block loop, id translation and 2-way merge
                                 (reference drivers/run_convdr_inference.py:157-242)
  * `write_block` / `read_block` — the per-rank pickle block format
                                 (reference utils/util.py:108-111, drivers/gen_passage_embeddings.py:156-167)
  * `truth_fp64`, `compare`    — fp64 ground truth and the parity comparator of BASELINE.json
                                 (ids equal except at ties within 1e-5 relative, scores within 1e-5 relative)

PARITY STATUS: **unpinned for the FAISS arithmetic.**  The arithmetic of the path lives in the
third-party `faiss-gpu` wheel (requirements.txt:4, unpinned; contemporaneous releases 1.6.5-1.7.2),
which is absent from /root/reference and from this image, and the reference ships no tests, fixtures
or golden vectors.  `IndexFlatIP.search` therefore restates the published
algorithm of FAISS 1.7.x `knn_inner_product` (utils/distances.cpp): fp32 scores (per-query SIMD dot
for nq < 20, blocked sgemm 4096 x 1024 otherwise), a min-heap that inserts on strict `>` (so at the
k-th boundary the lower index wins), results reordered descending, `-1` / `-FLT_MAX` padding when
ntotal < k.  The order *within* a group of exactly equal scores is FAISS-version dependent; this
oracle uses (score desc, index asc) and `compare` tolerates permutations inside tie groups.
What IS pinned: `search_one_by_one` is checked against the reference's own function, imported from
/root/reference with stub modules (tests/golden/make_golden.py -> tests/golden/*.npz).
"""
from __future__ import annotations

import os
import pickle
from typing import Optional

import numpy as np

FLT_MAX = np.float32(3.4028234663852886e38)
BLAS_THRESHOLD = 20      # faiss::distance_compute_blas_threshold
BS_QUERY = 4096          # faiss::distance_compute_blas_query_bs
BS_DB = 1024             # faiss::distance_compute_blas_database_bs


def _gemm_f32(a: np.ndarray, b: np.ndarray, backend: str) -> np.ndarray:
    """fp32 a @ b.T through the host BLAS (OpenBLAS via numpy, or MKL via torch's CPU matmul)."""
    if backend == "torch":
        import torch
        return torch.matmul(torch.from_numpy(a), torch.from_numpy(b).T).numpy()
    return a @ b.T


def _block_topk(s_blk: np.ndarray, k: int, backend: str = "numpy"):
    """Per-row top-k of one score block with the oracle's tie rule (score desc, index asc).
    Fast path: argpartition; rows whose k-th value is tied with excluded entries are redone with a
    full stable sort (argpartition picks arbitrarily among equal values)."""
    nq, nb = s_blk.shape
    if nb <= k:
        order = np.argsort(-s_blk, axis=1, kind="stable")
        return np.take_along_axis(s_blk, order, axis=1), order.astype(np.int64)
    if backend == "torch":
        import torch
        tv, ti = torch.topk(torch.from_numpy(s_blk), k, dim=1, sorted=False)
        vals, part = tv.numpy(), ti.numpy()
    else:
        part = np.argpartition(-s_blk, k - 1, axis=1)[:, :k]
        vals = np.take_along_axis(s_blk, part, axis=1)
    thr = vals.min(axis=1, keepdims=True)
    tied_rows = np.nonzero((s_blk == thr).sum(axis=1) != (vals == thr).sum(axis=1))[0]
    order = np.lexsort((part, -vals), axis=1)          # score desc, then index asc
    vals = np.take_along_axis(vals, order, axis=1)
    part = np.take_along_axis(part, order, axis=1)
    for r in tied_rows:
        o = np.argsort(-s_blk[r], kind="stable")[:k]
        vals[r], part[r] = s_blk[r, o], o
    return vals, part.astype(np.int64)


def _topk_merge(best_s, best_i, s_blk, i0, k, backend: str = "numpy"):
    """Fold one score block [nq, nb] (db indices i0..i0+nb-1) into the running top-k.
    Order: score desc, index asc — equivalent to the strict-`>` heap at the k-th boundary."""
    blk_s, blk_i = _block_topk(s_blk, k, backend)
    blk_i = blk_i + i0
    if best_s is None:
        return blk_s, blk_i
    cat_s = np.concatenate([best_s, blk_s], axis=1)
    cat_i = np.concatenate([best_i, blk_i], axis=1)
    kk = min(k, cat_s.shape[1])
    # earlier blocks hold lower indices and come first, so a stable sort on -score keeps the
    # lower index first among equal scores
    order = np.argsort(-cat_s, axis=1, kind="stable")[:, :kk]
    return np.take_along_axis(cat_s, order, axis=1), np.take_along_axis(cat_i, order, axis=1)


def knn_inner_product(x: np.ndarray, xb: np.ndarray, k: int, backend: str = "numpy",
                      db_block: Optional[int] = None):
    """FAISS `knn_inner_product` restated: (D float32 [nq,k] desc, I int64 [nq,k])."""
    x = np.ascontiguousarray(x, dtype=np.float32)
    xb = np.ascontiguousarray(xb, dtype=np.float32)
    nq, n = x.shape[0], xb.shape[0]
    D = np.full((nq, k), -FLT_MAX, dtype=np.float32)
    I = np.full((nq, k), -1, dtype=np.int64)
    if nq == 0 or n == 0:
        return D, I
    bs_db = db_block or (BS_DB if nq >= BLAS_THRESHOLD else 65536)
    for q0 in range(0, nq, BS_QUERY):
        q1 = min(nq, q0 + BS_QUERY)
        best_s = best_i = None
        for j0 in range(0, n, bs_db):
            j1 = min(n, j0 + bs_db)
            s_blk = _gemm_f32(x[q0:q1], xb[j0:j1], backend)
            best_s, best_i = _topk_merge(best_s, best_i, s_blk, j0, k, backend)
        kk = best_s.shape[1]
        D[q0:q1, :kk] = best_s
        I[q0:q1, :kk] = best_i
    return D, I


class IndexFlatIP:
    """CPU stand-in for faiss.IndexFlatIP(d) — same surface as the product facade."""

    def __init__(self, d: int, backend: str = "numpy", db_block: Optional[int] = None):
        self.d = int(d)
        self.is_trained = True
        self.backend = backend
        self.db_block = db_block
        self._blocks: list[np.ndarray] = []
        self._xb: Optional[np.ndarray] = None

    @property
    def ntotal(self) -> int:
        return sum(b.shape[0] for b in self._blocks)

    def add(self, x) -> None:
        x = np.ascontiguousarray(x, dtype=np.float32)
        assert x.ndim == 2 and x.shape[1] == self.d
        self._blocks.append(x.copy())   # FAISS copies (the driver deletes its array, :203-204)
        self._xb = None

    def reset(self) -> None:
        self._blocks = []
        self._xb = None

    def _all(self) -> np.ndarray:
        if self._xb is None:
            self._xb = (np.concatenate(self._blocks, axis=0) if self._blocks
                        else np.zeros((0, self.d), dtype=np.float32))
        return self._xb

    def search(self, x, k: int):
        x = np.ascontiguousarray(x, dtype=np.float32)
        assert x.ndim == 2 and x.shape[1] == self.d
        return knn_inner_product(x, self._all(), int(k), self.backend, self.db_block)


# ---------------------------------------------------------------------------------------------
# Block format: reference utils/util.py:108-111 pickles (protocol 4) the rank's float32 [n,768]
# embedding array and its int64 [n] offset array to "{prefix}_data_obj_{rank}.pb" with prefixes
# "passage__emb_p_" / "passage__embid_p_" (gen_passage_embeddings.py:158-166); the reader side is
# run_convdr_inference.py:161-175.
# ---------------------------------------------------------------------------------------------
def block_paths(ann_data_dir: str, block_id: int):
    return (os.path.join(ann_data_dir, "passage__emb_p__data_obj_%d.pb" % block_id),
            os.path.join(ann_data_dir, "passage__embid_p__data_obj_%d.pb" % block_id))


def write_block(ann_data_dir: str, block_id: int, emb: np.ndarray, embid: np.ndarray) -> None:
    os.makedirs(ann_data_dir, exist_ok=True)
    pe, pi = block_paths(ann_data_dir, block_id)
    with open(pe, "wb") as f:
        pickle.dump(np.ascontiguousarray(emb, dtype=np.float32), f, protocol=4)
    with open(pi, "wb") as f:
        pickle.dump(np.ascontiguousarray(embid, dtype=np.int64), f, protocol=4)


def read_block(ann_data_dir: str, block_id: int):
    pe, pi = block_paths(ann_data_dir, block_id)
    with open(pe, "rb") as f:
        emb = pickle.load(f)
    with open(pi, "rb") as f:
        embid = pickle.load(f)
    return emb, embid


def search_one_by_one(ann_data_dir: str, index, query_embedding: np.ndarray, topN: int, max_blocks: int = 8):
    """Restatement of reference run_convdr_inference.py:157-242 (array form of its list logic).

    For block 0..7 (stop at the first block that cannot be loaded, :176-177): add, search topN,
    translate local indices through the block's offset array (numpy fancy indexing, so `-1`
    padding wraps to the block's LAST offset, :190), reset; then 2-way merge the first topN entries
    of the running list with the block's topN, `>=` keeping the running entry on ties (:218), both
    tails drained (:224-229) — the result is 2*topN wide once two blocks were merged.  Returns
    (merged_D float64, merged_I int64), as `np.array` of Python floats / ints gives (:231-238).
    """
    run_s = run_i = None
    for block_id in range(max_blocks):
        try:
            emb, embid = read_block(ann_data_dir, block_id)
        except Exception:          # the reference uses a bare except -> break
            break
        index.add(emb)
        D, I = index.search(query_embedding, topN)
        ids = np.asarray(embid)[I]
        index.reset()
        cur_s = D.astype(np.float64)        # D.tolist() -> Python floats
        cur_i = ids.astype(np.int64)
        if run_s is None:
            run_s, run_i = cur_s, cur_i
            continue
        a_s, a_i = run_s[:, :topN], run_i[:, :topN]
        cat_s = np.concatenate([a_s, cur_s[:, :topN]], axis=1)
        cat_i = np.concatenate([a_i, cur_i[:, :topN]], axis=1)
        order = np.argsort(-cat_s, axis=1, kind="stable")   # stable: running list first on ties
        run_s = np.take_along_axis(cat_s, order, axis=1)
        run_i = np.take_along_axis(cat_i, order, axis=1)
    if run_s is None:
        raise TypeError("'NoneType' object is not iterable")   # what the reference raises with no block
    return run_s, run_i


def eval_rank_dedup(merged_D, merged_I, topN: int, offset2pid):
    """What `EvalDevQuery` does with the merged ranking before anything is written (reference
    drivers/run_convdr_inference.py:37-69): per query, the first topN (offset, score) pairs in rank order,
    `pred_pid = offset2pid[idx]` (:62; a negative idx wraps like any Python sequence index), a pid already in
    `seen_pid` is skipped (:64), the others take ranks 0, 1, ... (:65-71); slots never reached keep the
    pre-filled `(0, 0)` (:50).  Returns (pids int64 [nq, topN], scores float64 [nq, topN], counts int32 [nq])."""
    nq = len(merged_I)
    pids = np.zeros((nq, topN), dtype=np.int64)
    scores = np.zeros((nq, topN), dtype=np.float64)
    counts = np.zeros((nq,), dtype=np.int32)
    for query_idx in range(nq):
        seen_pid = set()
        rank = 0
        selected_ann_idx = merged_I[query_idx][:topN]
        selected_ann_score = np.asarray(merged_D[query_idx][:topN]).tolist()
        for idx, score in zip(selected_ann_idx, selected_ann_score):
            pred_pid = int(offset2pid[int(idx)])
            if pred_pid not in seen_pid:
                pids[query_idx, rank] = pred_pid
                scores[query_idx, rank] = score
                rank += 1
                seen_pid.add(pred_pid)
        counts[query_idx] = rank
    return pids, scores, counts


# ---------------------------------------------------------------------------------------------
# Ground truth and comparator
# ---------------------------------------------------------------------------------------------
def truth_fp64(x: np.ndarray, xb: np.ndarray, k: int, block: int = 262144):
    """Exact top-k with float64 scores: (D float64 [nq,k], I int64 [nq,k]), order (score desc, index asc)."""
    x64 = np.asarray(x, dtype=np.float64)
    nq, n = x64.shape[0], xb.shape[0]
    best_s = best_i = None
    for j0 in range(0, n, block):
        j1 = min(n, j0 + block)
        s = x64 @ np.asarray(xb[j0:j1], dtype=np.float64).T
        best_s, best_i = _topk_merge(best_s, best_i, s, j0, k)
    kk = 0 if best_s is None else best_s.shape[1]
    D = np.full((nq, k), -np.inf)
    I = np.full((nq, k), -1, dtype=np.int64)
    if kk:
        D[:, :kk], I[:, :kk] = best_s, best_i
    return D, I


def compare(D, I, D_ref, I_ref, score_of=None, rtol: float = 1e-5, atol: float = 0.0):
    """Parity comparator (BASELINE.json north_star): ids must match the reference position by
    position except where the two ids involved have scores within `rtol` relative (a tie: swap
    inside a tie group, or exchange across the k-th boundary with a tied (k+1)-th); scores must agree
    within `rtol` relative.  `score_of(q, ids) -> float64 scores` supplies exact scores for ids that
    appear in only one of the two lists (needed for boundary exchanges).  `atol` is an absolute
    floor for collections so small that near-zero scores are returned (a relative bound on a score
    of 1e-4 is below the fp32 rounding of the reference itself); 0 for real top-k regimes.

    Returns dict(exact_rows, tie_excused, violations, max_rel_score_err).
    """
    D = np.asarray(D, dtype=np.float64)
    D_ref = np.asarray(D_ref, dtype=np.float64)
    I = np.asarray(I)
    I_ref = np.asarray(I_ref)
    assert D.shape == D_ref.shape == I.shape == I_ref.shape
    nq, k = I.shape
    exact_rows = tie_excused = violations = 0
    max_rel = 0.0
    for q in range(nq):
        valid = I_ref[q] >= 0
        denom = np.maximum(np.abs(D_ref[q][valid]), 1e-30)
        if valid.any():
            err = np.abs(D[q][valid] - D_ref[q][valid])
            rel = err / denom
            bad = err > rtol * np.abs(D_ref[q][valid]) + atol
            max_rel = max(max_rel, float(rel[~bad].max()) if (~bad).any() else 0.0, float(rel[bad].max()) if bad.any() else 0.0)
            violations += int(bad.sum())
        if not np.array_equal(valid, I[q] >= 0):
            violations += 1
            continue
        if np.array_equal(I[q], I_ref[q]):
            exact_rows += 1
            continue
        ref_score = {int(i): float(s) for i, s in zip(I_ref[q], D_ref[q])}
        for pos in np.nonzero(I[q] != I_ref[q])[0]:
            a, b = int(I[q, pos]), int(I_ref[q, pos])
            if score_of is not None:
                sa, sb = score_of(q, np.array([a, b], dtype=np.int64))
            else:  # best effort without exact scores: the reference's own score for `a` if it lists it
                sa, sb = ref_score.get(a, D[q, pos]), D_ref[q, pos]
            if abs(sa - sb) <= rtol * max(abs(sa), abs(sb), 1e-30) + atol:
                tie_excused += 1
            else:
                violations += 1
    return dict(exact_rows=exact_rows, tie_excused=tie_excused, violations=violations,
                max_rel_score_err=max_rel)
