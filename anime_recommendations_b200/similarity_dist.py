"""Multi-GPU cosine top-k (BASELINE cfg3): one process per GPU, the fp32 table replicated.

* shard="candidates" (the north-star's scheme): rank r owns candidate rows [lo_r, hi_r); every rank runs the
  tensor-core candidate pass + exact fp32 re-rank of ALL queries against its own rows, the per-rank top-k
  lists (global row ids) are all-gathered with NCCL on the library's communicator and merged on the GPU
  (ar_topk_merge).  Each shard's list is the exact top-k of that shard, so the merge is the exact global top-k.
* shard="queries": rank r answers query rows [lo_r, hi_r) against the whole table -- no data-path
  collective at all (the bf16 table is 90 MB); the all-gather at the end only assembles the result.
* score_topk_sharded (BASELINE cfg4, model_recs.py:373-456 over many users): the query USERS are sharded, the anime
  table and the head are replicated, every rank scores its users independently -- no collective on the data
  path; `gather=True` assembles the result on every rank.
"""
from __future__ import annotations

import torch

from . import similarity as sim


def shard_range(n, rank, world, align=1):
    """Contiguous, balanced [lo, hi) of rank `rank` over n rows; `align` keeps boundaries on tile multiples."""
    units = (n + align - 1) // align
    base, rem = divmod(units, world)
    lo_u = rank * base + min(rank, rem)
    hi_u = lo_u + base + (1 if rank < rem else 0)
    return min(n, lo_u * align), min(n, hi_u * align)


def allpairs_topk_sharded(W, k, comm, kprime=16, shard="candidates", stats=None):
    """Exact top-k most cosine-similar OTHER rows for every row of W, computed by `comm.world` GPUs.
    -> (idx [n,k] int32, score [n,k] float32) on every rank."""
    W = sim.as_table(W)
    n = W.shape[0]
    rank, world = comm.rank, comm.world
    Wn, res = sim.normalize_rows_bf16(W, with_resid=True)
    res = torch.nan_to_num(res, nan=1.0)
    rmax = float(res.max().item())
    if shard == "candidates":
        lo, hi = shard_range(n, rank, world, align=128)
        words = (n + 31) // 32
        smask = None

        def exact_row(r):
            nonlocal smask
            if smask is None:                                   # bit mask of this rank's candidate rows
                m = torch.zeros(n, dtype=torch.bool, device=W.device)
                m[lo:hi] = True
                smask = sim.pack_mask(m.cpu().numpy(), n, W.device)
            return sim.cosine_topk_query_device(W, r, k, mask_bits=smask, exclude=r)

        st = {} if stats is None else stats
        oi, os_ = sim._certified_topk(W, Wn, res, W, Wn, rmax, k, kprime, True, None, None, st, exact_row,
                                      c_range=(lo, hi))
        gi, gs = comm.allgather(oi), comm.allgather(os_)         # [world, n, k]
        return sim.topk_merge(gi, gs, k)
    if shard == "queries":
        lo, hi = shard_range(n, rank, world, align=256)
        oi, os_ = sim.allpairs_topk(W, k=k, kprime=kprime, q0=lo, nq=hi - lo, stats=stats)
        # ragged shards: pad to the largest, gather, cut
        cap = max(shard_range(n, r, world, align=256)[1] - shard_range(n, r, world, align=256)[0] for r in range(world))
        pi = torch.full((cap, k), -1, dtype=torch.int32, device=W.device)
        ps = torch.full((cap, k), float("-inf"), dtype=torch.float32, device=W.device)
        pi[:hi - lo], ps[:hi - lo] = oi, os_
        gi, gs = comm.allgather(pi), comm.allgather(ps)
        parts_i, parts_s = [], []
        for r in range(world):
            a, b = shard_range(n, r, world, align=256)
            parts_i.append(gi[r, :b - a])
            parts_s.append(gs[r, :b - a])
        return torch.cat(parts_i), torch.cat(parts_s)
    raise ValueError("shard must be 'candidates' or 'queries'")


def shard_csr(indptr, idx, lo, hi):
    """Rows [lo, hi) of a host CSR (watched lists), re-based to start at 0."""
    import numpy as np
    indptr = np.asarray(indptr, np.int64)
    b, e = int(indptr[lo]), int(indptr[hi])
    return (indptr[lo:hi + 1] - b).astype(np.int64), np.asarray(idx)[b:e]


def score_topk_sharded(model, users, watched_indptr, watched_idx, k, rank, world, comm=None, gather=False,
                       cand_mask=None, stats=None):
    """similarity.score_topk with the query users split evenly over `world` ranks (rank r takes the contiguous
    slice shard_range(len(users), r, world)).  Returns this rank's (lo, hi, idx, prediction); with gather=True and
    a `comm` (dist.Comm) every rank gets the full (idx [n,k], prediction [n,k]) instead."""
    import numpy as np
    users = np.asarray(users)
    n = len(users)
    lo, hi = shard_range(n, rank, world)
    ip, ix = shard_csr(watched_indptr, watched_idx, lo, hi)
    if hi > lo:
        oi, pr = sim.score_topk(model, users[lo:hi], ip, ix, k, cand_mask=cand_mask, stats=stats)
    else:
        oi, pr = np.zeros((0, k), np.int32), np.zeros((0, k), np.float32)
    if not gather:
        return lo, hi, oi, pr
    if comm is None:
        raise ValueError("gather=True needs a communicator")
    cap = max(shard_range(n, r, world)[1] - shard_range(n, r, world)[0] for r in range(world))
    dev = model.device
    pi = torch.full((cap, k), -1, dtype=torch.int32, device=dev)
    ps = torch.full((cap, k), float("-inf"), dtype=torch.float32, device=dev)
    pi[:hi - lo], ps[:hi - lo] = torch.from_numpy(oi).to(dev), torch.from_numpy(pr).to(dev)
    gi, gs = comm.allgather(pi), comm.allgather(ps)
    parts_i, parts_s = [], []
    for r in range(world):
        a, b = shard_range(n, r, world)
        parts_i.append(gi[r, :b - a])
        parts_s.append(gs[r, :b - a])
    return torch.cat(parts_i).cpu().numpy(), torch.cat(parts_s).cpu().numpy()
