"""Multi-GPU cosine top-k (BASELINE cfg3): one process per GPU, the fp32 table replicated.

* shard="candidates" (the north-star's scheme): rank r owns candidate rows [lo_r, hi_r); every rank runs the
  tensor-core candidate pass + exact fp32 re-rank of ALL queries against its own rows, the per-rank top-k
  lists (global row ids) are all-gathered with NCCL on the library's communicator and merged on the GPU
  (ar_topk_merge).  Each shard's list is the exact top-k of that shard, so the merge is the exact global top-k.
* shard="queries": rank r answers query rows [lo_r, hi_r) against the whole table -- no data-path
  collective at all (the bf16 table is 90 MB); the all-gather at the end only assembles the result.
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
