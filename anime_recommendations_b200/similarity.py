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


BF16_SCORE_EPS = 0.004   # |bf16-operand score - fp32 score| <= 2^-8 for unit vectors (+ accumulation slack)


def normalize_rows_bf16(W):
    """Row-normalised bf16 copy of a table: the operand format of the tensor-core pass."""
    W = as_table(W)
    out = torch.empty(W.shape, dtype=torch.bfloat16, device=W.device)
    check(lib().ar_rownorm_bf16(ptr(W), W.shape[0], W.shape[1], ptr(out), stream_ptr()), "ar_rownorm_bf16")
    return out


def allpairs_candidates(Qn, q0, nq, Cn, c0, nc, kprime=16, exclude_self=False, watched=None, dump=False):
    """Tensor-core candidate pass: per query row and candidate chunk the kprime best rows by bf16 score.
    -> (idx [n_chunks, nq, kprime] int32, score same float32, dump [nq, nc] or None), device tensors."""
    L = lib()
    n_chunks = int(L.ar_allpairs_chunks(nq, nc))
    dev = Qn.device
    oi = torch.empty((n_chunks, nq, kprime), dtype=torch.int32, device=dev)
    os_ = torch.empty((n_chunks, nq, kprime), dtype=torch.float32, device=dev)
    dm = torch.zeros((nq, nc), dtype=torch.float32, device=dev) if dump else None
    stride = 0 if watched is None else watched.shape[1]
    check(L.ar_cosine_topk_allpairs(ptr(Qn), Qn.shape[0], q0, nq, ptr(Cn), Cn.shape[0], c0, nc, Qn.shape[1], kprime,
                                    1 if exclude_self else 0, ptr(watched), stride, n_chunks, ptr(oi), ptr(os_),
                                    ptr(dm), stream_ptr()), "ar_cosine_topk_allpairs")
    return oi, os_, dm


def rerank(Wq, q0, nq, Wc, cand, k, cand_score=None, eps=BF16_SCORE_EPS):
    """Exact fp32 cosine re-rank of candidate lists cand [n_lists, nq, list_k] (or [nq, list_k]).
    -> (idx [nq,k], score [nq,k], certified [nq] uint8 or None), device tensors."""
    if cand.dim() == 2:
        cand = cand.unsqueeze(0)
        cand_score = None if cand_score is None else cand_score.unsqueeze(0)
    nl, _, lk = cand.shape
    oi = torch.empty((nq, k), dtype=torch.int32, device=Wq.device)
    os_ = torch.empty((nq, k), dtype=torch.float32, device=Wq.device)
    cert = torch.empty(nq, dtype=torch.uint8, device=Wq.device) if cand_score is not None else None
    check(lib().ar_cosine_rerank(ptr(Wq), q0, nq, ptr(Wc), Wq.shape[1], ptr(cand.contiguous()),
                                 ptr(None if cand_score is None else cand_score.contiguous()), nl, lk, k, eps,
                                 ptr(oi), ptr(os_), ptr(cert), stream_ptr()), "ar_cosine_rerank")
    return oi, os_, cert


def allpairs_topk(W, k=10, kprime=16, q0=0, nq=None, Wn_bf16=None, stats=None):
    """BASELINE cfg3: for every query row in [q0, q0+nq) the k most cosine-similar OTHER rows of W.

    bf16 tensor-core pass selects kprime candidates per row (per candidate chunk), the fp32 re-rank orders
    them exactly, and rows whose result cannot be certified exact (bf16 error bound) fall back to the fp32
    single-query kernel.  -> (idx [nq,k] int32, score [nq,k] float32) device tensors."""
    W = as_table(W)
    n = W.shape[0]
    nq = n - q0 if nq is None else nq
    Wn = normalize_rows_bf16(W) if Wn_bf16 is None else Wn_bf16
    ci, cs, _ = allpairs_candidates(Wn, q0, nq, Wn, 0, n, kprime, exclude_self=True)
    oi, os_, cert = rerank(W, q0, nq, W, ci, k, cand_score=cs)
    bad = torch.nonzero(cert == 0).reshape(-1).cpu().tolist()
    for r in bad:                                      # exact fp32 path for the uncertified few
        fi, fs = cosine_topk_query_device(W, q0 + r, k, exclude=q0 + r)
        oi[r], os_[r] = fi, fs
    if stats is not None:
        stats.update(uncertified=len(bad), n_chunks=ci.shape[0], kprime=kprime)
    return oi, os_


def watched_bits(indptr, idx, n_rows, n_cols, device, cand_mask=None):
    """[n_rows, ceil(n_cols/32)] uint32 rows (as int32 tensor): bit set = candidate is dropped."""
    words = (n_cols + 31) // 32
    if cand_mask is None:
        out = torch.zeros((n_rows, words), dtype=torch.int32, device=device)
    else:
        inv = pack_mask(~np.asarray(cand_mask, dtype=bool), n_cols, device)
        out = inv.reshape(1, words).repeat(n_rows, 1).contiguous()
    ip = torch.as_tensor(np.asarray(indptr, np.int64)).to(device)
    ix = torch.as_tensor(np.asarray(idx, np.int32)).to(device)
    check(lib().ar_bits_from_csr(ptr(ip), ptr(ix), n_rows, words, n_cols, ptr(out), stream_ptr()), "ar_bits_from_csr")
    return out


def score_topk(model, users, watched_indptr, watched_idx, k, cand_mask=None, kprime=32, stats=None):
    """model_recs over many users (BASELINE cfg4): predicted rating of every anime the user has NOT rated,
    top-k by Prediction (model_recs.py:132-192,373-456).  Prediction is a monotone map of cos(u, a)
    (Dense(1) -> BatchNorm(inference) -> sigmoid), so candidates are ranked by sign(w*gamma)*cos on the
    tensor cores, re-ranked in fp32, and only the winners go through the exact forward (ar_predict).
    -> (idx [n,k] int32 numpy, prediction [n,k] float32 numpy); -1 / -inf pad short lists."""
    model._sync_tables()
    dev = model.device
    users_t = torch.as_tensor(np.asarray(users, np.int64)).to(dev)
    nq, na = users_t.numel(), model.n_anime
    head = model.head.cpu().numpy()
    sign = -1.0 if float(head[0]) * float(head[2]) < 0 else 1.0
    Uq = (model.U[users_t] * sign).contiguous()               # query rows (negated when the map decreases)
    Qn, Cn = normalize_rows_bf16(Uq), normalize_rows_bf16(model.A)
    wb = watched_bits(watched_indptr, watched_idx, nq, na, dev, cand_mask)
    ci, cs, _ = allpairs_candidates(Qn, 0, nq, Cn, 0, na, kprime, exclude_self=False, watched=wb)
    oi, os_, cert = rerank(Uq, 0, nq, model.A, ci, k, cand_score=cs)
    bad = torch.nonzero(cert == 0).reshape(-1).cpu().tolist()
    for r in bad:
        fi, fs = _query_vs_table(Uq[r], model.A, k, wb[r])
        oi[r], os_[r] = fi, fs
    valid = oi >= 0
    iu = users_t.to(torch.int32).reshape(-1, 1).expand(-1, k)[valid].contiguous()
    ia = oi[valid].contiguous()
    pred = torch.full((nq, k), float("-inf"), dtype=torch.float32, device=dev)
    if iu.numel():
        out = torch.empty(iu.numel(), dtype=torch.float32, device=dev)
        check(lib().ar_predict(ptr(model.U), ptr(model.A), model.dim, ptr(model.head), ptr(model.bn_moving), ptr(iu),
                               ptr(ia), iu.numel(), ptr(out), stream_ptr()), "ar_predict")
        pred[valid] = out
    if stats is not None:
        stats.update(uncertified=len(bad), n_chunks=ci.shape[0], kprime=kprime, sign=sign)
    return oi.cpu().numpy(), pred.cpu().numpy()


def _query_vs_table(qrow, C, k, drop_bits):
    """fp32 fallback for one scoring row: append the query to the table and use the single-query kernel."""
    T = torch.cat([C, qrow.reshape(1, -1)], dim=0).contiguous()
    n = C.shape[0]
    words = (n + 1 + 31) // 32
    keep = torch.zeros(words, dtype=torch.int32, device=C.device)
    keep[:drop_bits.numel()] = ~drop_bits
    return cosine_topk_query_device(T, n, k, mask_bits=keep, exclude=n)


def topk_merge(idx, score, k_out):
    """Merge [n_lists, n_queries, k_in] partial lists (global row ids) into [n_queries, k_out]."""
    nl, nq, kin = idx.shape
    oi = torch.empty((nq, k_out), dtype=torch.int32, device=idx.device)
    os_ = torch.empty((nq, k_out), dtype=torch.float32, device=idx.device)
    check(lib().ar_topk_merge(ptr(idx), ptr(score), nl, nq, kin, k_out, ptr(oi), ptr(os_), stream_ptr()),
          "ar_topk_merge")
    return oi, os_
