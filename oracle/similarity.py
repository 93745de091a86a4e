"""Oracle for Half B: cosine-similarity top-k (TEST INFRASTRUCTURE).

The reference's similarity code is plain NumPy, so these functions are the
reference's own expressions (not a re-derivation):

* get_weights      similar_anime/similar_anime.py:159-170 (== similar_users.py:89-100,
                   user_recs/user_recs.py:182-193, helper_functions/load.py:35-41,
                   neural_network/neural_network.py:135-137)
* query_scores     similar_anime.py:404, similar_users.py:293, user_recs.py:475
* similar_users    similar_users.py:293-296,300-312
* similar_anime    similar_anime.py:408-409,438-468
* model_scores     model_recs/model_recs.py:394 (Keras predict, inference mode)

Tie / NaN policy (documented deviation): the reference ranks with
NumPy's unstable introsort and pandas' default quicksort, so the order among
EQUAL float32 scores is unspecified there.  Oracle and library both define it
as "higher score first, then lower row index"; NaN scores (zero-norm rows, no
epsilon in get_weights) are treated as -inf, i.e. never recommended.
"""
from __future__ import annotations

import numpy as np

from . import train as _train


def get_weights(W):
    """Row-normalise a table exactly as the reference does (no epsilon)."""
    W = np.asarray(W, dtype=np.float32)
    return W / np.linalg.norm(W, axis=1).reshape((-1, 1))


def query_scores(Wn, q):
    """dists = np.dot(weights, weights[encoded_index])."""
    return np.dot(Wn, Wn[q])


def rank_desc(scores, k, mask=None, exclude=None):
    """Top-k under the documented order.  mask: bool (n,), True = candidate."""
    s = np.asarray(scores, dtype=np.float32).copy()
    s[np.isnan(s)] = -np.inf
    if mask is not None:
        s[~np.asarray(mask, dtype=bool)] = -np.inf
    if exclude is not None and exclude >= 0:
        s[exclude] = -np.inf
    order = np.lexsort((np.arange(len(s)), -s.astype(np.float64)))
    order = order[np.isfinite(s[order])][:k]
    return order.astype(np.int32), s[order]


def similar_users(W, q, n):
    """find_similar_users: top-(n+1) of all rows, drop the query, descending."""
    Wn = get_weights(W)
    dists = query_scores(Wn, q)
    idx, sc = rank_desc(dists, n + 1)
    keep = idx != q
    return idx[keep], sc[keep]


def similar_anime(W, q, count, mask=None):
    """anime_recs: rank all, keep `mask` (Type/genre filter), drop the query, first `count`."""
    Wn = get_weights(W)
    dists = query_scores(Wn, q)
    return rank_desc(dists, count, mask=mask, exclude=q)


def allpairs_topk(W, k, q0=0, q1=None, block=2048, exclude_self=True):
    """BASELINE cfg3: the single-query path looped over every query row (blocked GEMM)."""
    Wn = get_weights(W)
    n = Wn.shape[0]
    q1 = n if q1 is None else q1
    out_i = np.full((q1 - q0, k), -1, dtype=np.int32)
    out_s = np.full((q1 - q0, k), -np.inf, dtype=np.float32)
    for b0 in range(q0, q1, block):
        b1 = min(q1, b0 + block)
        S = Wn[b0:b1] @ Wn.T
        for r in range(b0, b1):
            i, s = rank_desc(S[r - b0], k, exclude=r if exclude_self else None)
            out_i[r - q0, :len(i)] = i
            out_s[r - q0, :len(s)] = s
    return out_i, out_s


def allpairs_topk_fast(W, k, q0=0, q1=None, block=1024, exclude_self=True):
    """Same result as allpairs_topk for tie-free data, argpartition-based (CPU baseline leg)."""
    Wn = get_weights(W)
    n = Wn.shape[0]
    q1 = n if q1 is None else q1
    out_i = np.empty((q1 - q0, k), dtype=np.int32)
    out_s = np.empty((q1 - q0, k), dtype=np.float32)
    for b0 in range(q0, q1, block):
        b1 = min(q1, b0 + block)
        S = Wn[b0:b1] @ Wn.T
        if exclude_self:
            S[np.arange(b1 - b0), np.arange(b0, b1)] = -np.inf
        part = np.argpartition(-S, k - 1, axis=1)[:, :k]
        ps = np.take_along_axis(S, part, axis=1)
        o = np.lexsort((part, -ps.astype(np.float64)), axis=1)
        out_i[b0 - q0:b1 - q0] = np.take_along_axis(part, o, axis=1)
        out_s[b0 - q0:b1 - q0] = np.take_along_axis(ps, o, axis=1)
    return out_i, out_s


def model_scores(st: "_train.State", user, anime_idx):
    """ratings = model.predict([user repeated, anime_idx]).flatten()."""
    iu = np.full(len(anime_idx), user, dtype=np.int64)
    return _train.predict(st, iu, anime_idx).reshape(-1)


def score_topk(st: "_train.State", users, watched_indptr, watched_idx, k, cand_mask=None):
    """model_recs over many users: predict every anime, drop watched, top-k by Prediction.

    watched CSR: watched_idx[watched_indptr[j]:watched_indptr[j+1]] are the anime
    indices users[j] has rated (model_recs.py:144-155).
    """
    na = st.A.shape[0]
    all_anime = np.arange(na, dtype=np.int64)
    out_i = np.full((len(users), k), -1, dtype=np.int32)
    out_p = np.full((len(users), k), -np.inf, dtype=np.float32)
    for j, u in enumerate(users):
        pred = model_scores(st, u, all_anime)
        mask = np.ones(na, dtype=bool) if cand_mask is None else np.asarray(cand_mask, dtype=bool).copy()
        mask[watched_idx[watched_indptr[j]:watched_indptr[j + 1]]] = False
        i, p = rank_desc(pred, k, mask=mask)
        out_i[j, :len(i)] = i
        out_p[j, :len(p)] = p
    return out_i, out_p
