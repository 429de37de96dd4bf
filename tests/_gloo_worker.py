"""Worker of tests/test_host_logic.py::test_sharded_search_over_gloo_world2_equals_single_index."""
import os

import numpy as np
import torch
import torch.distributed as dist

from convdr_b200 import synth
from convdr_b200.dist import ShardedFlatIP, shard_range
from oracle import flat_ip


def main():
    dist.init_process_group("gloo", init_method="env://")
    rank, world = dist.get_rank(), dist.get_world_size()
    P = synth.block(0, 5003, seed=11)
    Q = synth.block(0, 9, seed=11, stream=1)
    lo, hi = shard_range(P.shape[0], rank, world)
    local = flat_ip.IndexFlatIP(768)
    local.add(P[lo:hi])

    def local_search(q, k):
        D, I = local.search(q.numpy(), k)
        I = np.where(I >= 0, I + lo, -1)
        return torch.from_numpy(D), torch.from_numpy(I)

    def merge(Dp, Ip):  # stable merge, earlier shard first on ties — the contract of merge_kernel
        W, nq, k = Dp.shape
        cat_s = Dp.permute(1, 0, 2).reshape(nq, W * k).numpy()
        cat_i = Ip.permute(1, 0, 2).reshape(nq, W * k).numpy()
        order = np.argsort(-cat_s, axis=1, kind="stable")[:, :k]
        return (torch.from_numpy(np.take_along_axis(cat_s, order, 1)),
                torch.from_numpy(np.take_along_axis(cat_i, order, 1)))

    sh = ShardedFlatIP(index=None, local_search=local_search, merge=merge)
    D, I = sh.search(torch.from_numpy(Q), 25)
    np.savez(os.path.join(os.environ["OUT_DIR"], f"rank{rank}.npz"), D=D.numpy(), I=I.numpy())
    dist.barrier()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
