"""Collaborative aggregation of user_recs/user_recs.py:348-387,708-794 as a batched GPU job: favourites of every
user (ratings at or above the user's 80th percentile), then for every query user the anime most common among the
favourites of its similar users -- for ALL users at once from the all-pairs user top-k (similarity.allpairs_topk).
CUDA through the C-ABI (ar_user_favourites, ar_user_recs); there is no CPU path."""
from __future__ import annotations

import numpy as np
import torch

from ._capi import check, lib, ptr, stream_ptr


class RatingsCSR:
    """Ratings grouped by user row index (stable: file order inside a user), resident on the GPU."""

    def __init__(self, user_idx, anime_idx, rating, n_users, n_anime, device=None):
        dev = torch.device(device) if device is not None else torch.device("cuda", torch.cuda.current_device())
        user_idx = np.asarray(user_idx)
        order = np.argsort(user_idx, kind="stable")
        counts = np.bincount(user_idx, minlength=n_users)
        self.n_users, self.n_anime = int(n_users), int(n_anime)
        self.indptr_host = np.r_[0, np.cumsum(counts)].astype(np.int64)
        self.indptr = torch.from_numpy(self.indptr_host).to(dev)
        self.anime = torch.from_numpy(np.asarray(anime_idx)[order].astype(np.int32)).to(dev)
        self.rating = torch.from_numpy(np.asarray(rating)[order].astype(np.float32)).to(dev)
        self.order = order
        self.device = dev


def favourites(csr, percentile=80.0, return_thresholds=False):
    """uint8 flag per CSR entry: the rating is at or above its user's percentile (user_recs.py:359-361)."""
    fav = torch.empty(csr.rating.numel(), dtype=torch.uint8, device=csr.device)
    thr = torch.empty(csr.n_users, dtype=torch.float64, device=csr.device) if return_thresholds else None
    check(lib().ar_user_favourites(ptr(csr.indptr), ptr(csr.rating), csr.n_users, float(percentile), ptr(fav),
                                   ptr(thr) if thr is not None else None, stream_ptr()), "ar_user_favourites")
    return (fav, thr) if return_thresholds else fav


def similar_user_recs(csr, fav, query_users, sim_users, n_recs):
    """user_recs.py:708-794 for a batch of query users.  query_users: (Q,) row indices; sim_users: (Q, k) row
    indices of each query's similar users (< 0 = padding).  Returns (idx, cnt) int32 device tensors (Q, n_recs):
    anime row indices ranked by how many similar users hold them among their favourites (ties: lower index);
    idx = -1 past the last recommendation."""
    q = torch.as_tensor(query_users, dtype=torch.int32, device=csr.device).contiguous()
    s = torch.as_tensor(sim_users, dtype=torch.int32, device=csr.device).contiguous()
    if s.dim() != 2 or s.shape[0] != q.numel():
        raise ValueError("sim_users must be (n_query, k)")
    idx = torch.empty((q.numel(), int(n_recs)), dtype=torch.int32, device=csr.device)
    cnt = torch.empty_like(idx)
    check(lib().ar_user_recs(ptr(csr.indptr), ptr(csr.anime), ptr(fav), csr.n_anime, ptr(q), q.numel(), ptr(s),
                             s.shape[1], int(n_recs), ptr(idx), ptr(cnt), stream_ptr()), "ar_user_recs")
    return idx, cnt


def user_recs_all(model_user_table, csr, n_sim_users=10, n_recs=10, percentile=80.0):
    """Recommendations for EVERY user: all-pairs cosine top-k over the user table (tensor cores) -> favourites ->
    counts.  Returns (idx, cnt, sim_idx): (n_users, n_recs) x 2 and the (n_users, n_sim_users) similar users."""
    from . import similarity
    sim_idx, _ = similarity.allpairs_topk(model_user_table, k=int(n_sim_users))
    fav = favourites(csr, percentile)
    q = torch.arange(csr.n_users, dtype=torch.int32, device=csr.device)
    idx, cnt = similar_user_recs(csr, fav, q, sim_idx.to(torch.int32), n_recs)
    return idx, cnt, sim_idx
