"""Cosine-similarity top-k over embedding tables (Half B), host side.

Mirrors the NumPy code of similar_anime.py:136-171,399-468, similar_users.py:75-101,262-314,
user_recs.py:168-194,453-488 and the scoring of model_recs.py:373-456, with the arithmetic in
libanimerec.so.  Tables are torch CUDA tensors (or anything `as_table` can upload once).
"""
from __future__ import annotations

import ctypes as C

import numpy as np
import torch

from . import _capi
from ._capi import check, lib, ptr, stream_ptr


def as_table(W, device=None):
    """float32 (n, D) contiguous CUDA tensor view/copy of W."""
    if not torch.cuda.is_available():
        raise _capi.AnimerecError("similarity needs a CUDA device; there is no CPU fallback")
    dev = torch.device(device or "cuda:%d" % torch.cuda.current_device())
    if isinstance(W, torch.Tensor):
        return W.to(device=dev, dtype=torch.float32).contiguous()
    return torch.from_numpy(np.ascontiguousarray(W, np.float32)).to(dev)


def get_weights(model):
    """(anime_weights, user_weights) row-normalised float32 NumPy arrays -- similar_anime.py:136-171."""
    model._sync_tables()
    return normalize_rows(model.A).cpu().numpy(), normalize_rows(model.U).cpu().numpy()


def normalize_rows(W):
    W = as_table(W)
    out = torch.empty_like(W)
    check(lib().ar_rownorm(ptr(W), W.shape[0], W.shape[1], ptr(out), stream_ptr()), "ar_rownorm")
    return out


def pack_mask(mask, n, device):
    """bool (n,) -> uint32 bit words on device (bit r of word r>>5 set = row r is a candidate)."""
    if mask is None:
        return None
    m = np.asarray(mask, dtype=bool).reshape(-1)
    if m.shape[0] != n:
        raise ValueError("mask length %d != n_rows %d" % (m.shape[0], n))
    pad = (-n) % 32
    bits = np.packbits(np.concatenate([m, np.zeros(pad, bool)]), bitorder="little").view(np.uint32)
    return torch.from_numpy(bits.view(np.int32).copy()).to(device)


def cosine_topk_query(W, q, k, mask=None, exclude=None):
    """Top-k rows by cosine similarity to row q: (idx int32[<=k], score float32[<=k]), best first.

    `mask` restricts the candidates (Type/genre filter of similar_anime.py:438-463), `exclude`
    drops one row (the query itself, similar_anime.py:459)."""
    W = as_table(W)
    n, D = W.shape
    q = int(q)
    if not 0 <= q < n:
        raise KeyError("query row %d outside [0, %d)" % (q, n))
    if not 0 < k <= _capi.MAX_K:
        raise ValueError("k must be in [1, %d]" % _capi.MAX_K)
    L = lib()
    ws = torch.empty(max(8, L.ar_topk_query_workspace(n, k)), dtype=torch.uint8, device=W.device)
    oi = torch.empty(k, dtype=torch.int32, device=W.device)
    os_ = torch.empty(k, dtype=torch.float32, device=W.device)
    mbits = pack_mask(mask, n, W.device)
    check(L.ar_cosine_topk_query(ptr(W), n, D, q, ptr(mbits), -1 if exclude is None else int(exclude), k,
                                 ptr(oi), ptr(os_), ptr(ws), stream_ptr()), "ar_cosine_topk_query")
    oi, os_ = oi.cpu().numpy(), os_.cpu().numpy()
    keep = oi >= 0
    return oi[keep], os_[keep]


def cosine_topk_query_device(W, q, k, mask_bits=None, exclude=-1):
    """Same kernel, device tensors in and out (idx int32[k] with -1 padding, score float32[k]); no host copy."""
    n, D = W.shape
    L = lib()
    ws = torch.empty(max(8, L.ar_topk_query_workspace(n, k)), dtype=torch.uint8, device=W.device)
    oi = torch.empty(k, dtype=torch.int32, device=W.device)
    os_ = torch.empty(k, dtype=torch.float32, device=W.device)
    check(L.ar_cosine_topk_query(ptr(W), n, D, int(q), ptr(mask_bits), int(exclude), k, ptr(oi), ptr(os_), ptr(ws),
                                 stream_ptr()), "ar_cosine_topk_query")
    return oi, os_


def find_similar_users(W_users, q, n_users):
    """similar_users.py:293-312 / user_recs.py:475-488: top-(n+1) of ALL rows, then drop the query."""
    idx, sc = cosine_topk_query(W_users, q, n_users + 1)
    keep = idx != q
    return idx[keep], sc[keep]


def similar_anime(W_anime, q, count, mask=None):
    """similar_anime.py:404-468: rank all anime, keep the Type/genre candidates, drop the query."""
    return cosine_topk_query(W_anime, q, count, mask=mask, exclude=q)


def rerank(Wq, q0, nq, Wc, cand, k):
    """Exact fp32 cosine re-rank of per-query candidate lists (device tensors in, device tensors out)."""
    n_cand = cand.shape[1]
    oi = torch.empty((nq, k), dtype=torch.int32, device=Wq.device)
    os_ = torch.empty((nq, k), dtype=torch.float32, device=Wq.device)
    check(lib().ar_cosine_rerank(ptr(Wq), q0, nq, ptr(Wc), Wq.shape[1], ptr(cand), n_cand, k, ptr(oi), ptr(os_),
                                 stream_ptr()), "ar_cosine_rerank")
    return oi, os_


def topk_merge(idx, score, k_out):
    """Merge [n_lists, n_queries, k_in] partial lists (global row ids) into [n_queries, k_out]."""
    nl, nq, kin = idx.shape
    oi = torch.empty((nq, k_out), dtype=torch.int32, device=idx.device)
    os_ = torch.empty((nq, k_out), dtype=torch.float32, device=idx.device)
    check(lib().ar_topk_merge(ptr(idx), ptr(score), nl, nq, kin, k_out, ptr(oi), ptr(os_), stream_ptr()),
          "ar_topk_merge")
    return oi, os_
