"""One-process-per-GPU layout of the sharded index.

The collection is row-sharded across the ranks of a `torch.distributed` group (NCCL over
NVLink/NVSwitch on the GPU box, gloo in the CPU tests); the queries are replicated; every rank
computes its local top-k with global ids; ONE all-gather of the tiny `[nq, k]` (score, id) lists
follows, and a device k-way merge kernel produces the global top-k on every rank.

This replaces FAISS's `IndexShards` (host threads + PCIe + CPU merge) that the reference builds with
`index_cpu_to_gpu_multiple(..., shard=True)` (drivers/run_convdr_inference.py:355-368).
`torch.distributed` is plumbing only: the local search and the merge are libb2f kernels.
"""
from __future__ import annotations

from typing import Callable, Optional

import torch
import torch.distributed as dist


def shard_range(n_total: int, rank: int, world: int) -> tuple[int, int]:
    """Contiguous split, identical to b2f_add's per-device split (FAISS shard=True, successive ids)."""
    return n_total * rank // world, n_total * (rank + 1) // world


class ShardedFlatIP:
    """Rank-local view of a collection sharded over the process group.

    `local_search(q, k) -> (D [nq,k] float32, I [nq,k] int64 global ids)` and
    `merge(D_parts [W,nq,k], I_parts [W,nq,k]) -> (D, I)` default to the rank's FlatIPIndex; the CPU
    (gloo) tests inject stand-ins to exercise the partitioning and the collective plumbing.
    """

    def __init__(self, index=None, group=None, local_search: Optional[Callable] = None,
                 merge: Optional[Callable] = None):
        self.group = group
        self.rank = dist.get_rank(group) if dist.is_initialized() else 0
        self.world = dist.get_world_size(group) if dist.is_initialized() else 1
        self.index = index
        self._local_search = local_search or (lambda q, k: self.index.search_device(q, k))
        self._merge = merge or (lambda Dp, Ip: self.index.merge_device(Dp, Ip))
        self.ntotal = 0
        self._Dg = self._Ig = None

    # -- building -------------------------------------------------------------------------------
    def add_synthetic(self, n_total: int, seed: int = 0, stream: int = 0, norm: float = 1.0, chunk: int = 1 << 22):
        """Every rank generates its own slice of the synthetic stream on its GPU; ids are global rows."""
        lo, hi = shard_range(n_total, self.rank, self.world)
        self.index.reserve(hi - lo)
        for a in range(lo, hi, chunk):
            b = min(hi, a + chunk)
            self.index.add_synthetic(b - a, first_row=a, seed=seed, stream=stream, norm=norm, id_base=a)
        self.ntotal += n_total
        return lo, hi

    def add(self, x_global):
        """Every rank holds the same host array and keeps its slice (labels = global positions)."""
        import numpy as np
        n = x_global.shape[0]
        lo, hi = shard_range(n, self.rank, self.world)
        ids = np.arange(self.ntotal + lo, self.ntotal + hi, dtype=np.int64)
        self.index.add_with_ids(x_global[lo:hi], ids)
        self.ntotal += n
        return lo, hi

    # -- searching ------------------------------------------------------------------------------
    def search(self, q: torch.Tensor, k: int):
        """q: [nq, 768] float32 on this rank's device (replicated).  Returns the global (D, I)."""
        D, I = self._local_search(q, k)
        if self.world == 1:
            return D, I
        nq = q.shape[0]
        if self._Dg is None or self._Dg.shape != (self.world, nq, k) or self._Dg.device != D.device:
            self._Dg = torch.empty((self.world, nq, k), dtype=torch.float32, device=D.device)
            self._Ig = torch.empty((self.world, nq, k), dtype=torch.int64, device=D.device)
        # concatenated-along-dim-0 form: accepted by both NCCL and gloo
        dist.all_gather_into_tensor(self._Dg.view(self.world * nq, k), D.contiguous(), group=self.group)
        dist.all_gather_into_tensor(self._Ig.view(self.world * nq, k), I.contiguous(), group=self.group)
        if D.is_cuda:
            torch.cuda.current_stream(D.device).synchronize()  # the merge runs on the engine's stream
        return self._merge(self._Dg, self._Ig)

    def search_host(self, q_host, k: int, device=None):
        """End-to-end call with host buffers: pinned H2D of the queries, search, D2H of the result."""
        qt = torch.from_numpy(q_host)
        if device is not None:
            qt = qt.pin_memory().to(device, non_blocking=True)
        D, I = self.search(qt, k)
        return D.cpu().numpy(), I.cpu().numpy()
