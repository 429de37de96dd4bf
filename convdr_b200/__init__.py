"""convdr_b200 — B200-native exact inner-product top-k for ConvDR's dense-retrieval hot path.

Scope: the `faiss.IndexFlatIP` add / search / reset calls of
`drivers/run_convdr_inference.py::search_one_by_one`, the multi-GPU sharding that
`index_cpu_to_gpu_multiple(shard=True)` provided, and the per-block merge.

    import convdr_b200.faiss_compat as faiss     # the reference's `import faiss`
    index = faiss.IndexFlatIP(768); index.add(P); D, I = index.search(Q, 100); index.reset()

All compute is hand-written sm_100a CUDA in csrc/ behind the C ABI of include/b2f.h.
"""
from .index import FlatIPIndex, get_num_gpus, PATHS, DIM, MAX_K  # noqa: F401

__all__ = ["FlatIPIndex", "get_num_gpus", "PATHS", "DIM", "MAX_K"]
