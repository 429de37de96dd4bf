import numpy as np, torch, sys, os
sys.path.insert(0, "/root/repo")
from convdr_b200 import FlatIPIndex, synth
n = 60000
P = synth.block(0, n, seed=41)
P[50000:51500] = P[50000]
full = FlatIPIndex(768); full.add(P)
for lo, hi in ((0, 30000), (30000, 60000)):
    idx = FlatIPIndex(768)
    idx.add_with_ids(P[lo:hi], np.arange(lo, hi, dtype=np.int64))
    ref = FlatIPIndex(768); ref.set_option("path", "scan_exact")
    ref.add_with_ids(P[lo:hi], np.arange(lo, hi, dtype=np.int64))
    for nq, k, seed in ((1, 10, 1), (37, 100, 2), (173, 100, 3), (300, 64, 4), (8, 1000, 5)):
        qh = synth.block(0, nq, seed=seed, stream=1)
        q = torch.from_numpy(qh).cuda()
        D = torch.empty((nq, k), dtype=torch.float32, device="cuda"); I = torch.empty((nq, k), dtype=torch.int64, device="cuda")
        idx.reset_stats()
        idx.search_device_async(q, k, D, I)
        torch.cuda.synchronize()
        marked = (I[:, 0] == -2).sum().item()
        idx.finish()
        Dr, Ir = ref.search(qh, k)
        bad = (I.cpu().numpy() != Ir).sum()
        print(f"shard [{lo},{hi}) nq={nq} k={k}: marked={marked} fallback={idx.stat('fallback_queries')} mismatches={bad}", flush=True)
        if bad:
            rows = np.where((I.cpu().numpy() != Ir).any(axis=1))[0]
            print("  rows", rows[:10], "got", I.cpu().numpy()[rows[0]][:12], "want", Ir[rows[0]][:12])
