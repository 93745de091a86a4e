"""Host-side logic of the N > 1 paths on CPU: two `gloo` ranks (world_size 2, 127.0.0.1).

* candidate-sharded top-k: per-shard exact top-k lists, all-gathered and merged, equal the global top-k
  (the decomposition `similarity_dist.allpairs_topk_sharded` relies on; the lists here come from the oracle);
* shard_range tiles [0, n) exactly, on aligned boundaries;
* data-parallel batch layout: global step s = concatenation over ranks of each rank's s-th local batch.
"""
import os
import socket

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from anime_recommendations_b200.similarity_dist import shard_range
from oracle import similarity as osim


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, out):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        n, k = 700, 10
        W = np.random.RandomState(3).standard_normal((n, 32)).astype(np.float32)
        W[41] = W[17]                                         # exact tie across the two shards' boundary region
        Wn = osim.get_weights(W)
        lo, hi = shard_range(n, rank, world, align=128)
        li = np.full((n, k), -1, np.int32)
        ls = np.full((n, k), -np.inf, np.float32)
        mask = np.zeros(n, bool)
        mask[lo:hi] = True
        for q in range(n):
            i, s = osim.rank_desc(Wn @ Wn[q], k, mask=mask, exclude=q)
            li[q, :len(i)], ls[q, :len(i)] = i, s
        gi = [torch.empty((n, k), dtype=torch.int32) for _ in range(world)]
        gs = [torch.empty((n, k), dtype=torch.float32) for _ in range(world)]
        dist.all_gather(gi, torch.from_numpy(li))
        dist.all_gather(gs, torch.from_numpy(ls))
        ai, as_ = torch.stack(gi).numpy(), torch.stack(gs).numpy()
        ok = True
        for q in range(n):
            ci, cs = ai[:, q].reshape(-1), as_[:, q].reshape(-1)
            order = np.lexsort((ci, -cs))                     # score desc, then row id asc
            order = [o for o in order if ci[o] >= 0][:k]
            wi, _ = osim.rank_desc(Wn @ Wn[q], k, exclude=q)
            ok &= ci[order].tolist() == wi.tolist()
        # batch layout of the data-parallel trainer
        B, steps = 5, 4
        n_glob = world * B * steps - world * 2
        per = n_glob // world
        mine = torch.arange(rank * per, (rank + 1) * per)
        allr = [torch.empty(per, dtype=torch.int64) for _ in range(world)]
        dist.all_gather(allr, mine)
        order = np.concatenate([allr[r][s * B:(s + 1) * B].numpy() for s in range(steps) for r in range(world)])
        ok &= sorted(order.tolist()) == list(range(n_glob)) and order[:B].tolist() == list(range(B)) \
            and order[B:2 * B].tolist() == list(range(per, per + B))
        out[rank] = bool(ok)
    finally:
        dist.destroy_process_group()


def test_two_gloo_ranks_sharded_topk_merge_and_batch_layout():
    world = 2
    with mp.Manager() as mgr:
        out = mgr.dict()
        mp.spawn(_worker, args=(world, _free_port(), out), nprocs=world, join=True)
        assert dict(out) == {0: True, 1: True}


def test_shard_range_tiles_rows_on_aligned_boundaries():
    for n in (1, 127, 128, 129, 3001, 350000):
        for world in (1, 2, 3, 8):
            for align in (1, 128, 256):
                edges = [shard_range(n, r, world, align) for r in range(world)]
                assert edges[0][0] == 0 and edges[-1][1] == n
                for (a, b), (c, d) in zip(edges[:-1], edges[1:]):
                    assert b == c and a <= b and (b % align == 0 or b == n)


def test_shard_csr_tiles_the_watched_lists():
    """cfg4 scoring shards the query users: the per-rank CSR slices re-based to 0 tile the original lists."""
    from anime_recommendations_b200.similarity_dist import shard_csr
    rng = np.random.RandomState(0)
    n = 103
    counts = rng.randint(0, 9, n)
    indptr = np.r_[0, np.cumsum(counts)].astype(np.int64)
    idx = rng.randint(0, 50, indptr[-1]).astype(np.int32)
    for world in (1, 2, 3, 8):
        got_counts, got_idx = [], []
        for r in range(world):
            lo, hi = shard_range(n, r, world)
            ip, ix = shard_csr(indptr, idx, lo, hi)
            assert ip[0] == 0 and len(ip) == hi - lo + 1 and ip[-1] == len(ix)
            got_counts.append(np.diff(ip))
            got_idx.append(ix)
        np.testing.assert_array_equal(np.concatenate(got_counts), counts)
        np.testing.assert_array_equal(np.concatenate(got_idx), idx)


def test_replica_slices_rebuild_every_global_batch():
    """dist_fit.replica_slice (the data-parallel split of neural_network.py:176): the replicas' slices of a step, in
    rank order, are that step's global batch -- full steps and the shorter last one -- and every sample of the cut
    order is visited exactly once."""
    from anime_recommendations_b200.dist_fit import replica_slice
    rng = np.random.RandomState(0)
    for G, B, N in [(2, 5, 47), (4, 3, 36), (8, 4, 1000), (1, 7, 20), (3, 10, 29), (2, 8, 16), (4, 2, 3)]:
        order = rng.permutation(N)
        sl = [replica_slice(order, G, B, r) for r in range(G)]
        n_use = N // G * G
        assert all(len(s) == n_use // G for s in sl)
        assert sorted(np.concatenate(sl).tolist()) == sorted(order[:n_use].tolist())
        GB = G * B
        for s in range(-(-n_use // GB)):
            glob = order[s * GB:min(n_use, (s + 1) * GB)]
            nb = len(glob) // G
            got = np.concatenate([sl[r][s * B:s * B + nb] for r in range(G)])
            assert got.tolist() == glob.tolist(), (G, B, N, s)
