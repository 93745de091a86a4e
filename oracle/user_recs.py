"""CPU restatement (test infrastructure only) of the collaborative aggregation of user_recs/user_recs.py.

  favourites(u)  user_recs.py:359-361 (fave_genres) / 380-382 (fave_sources): the ratings of u at or above
                 np.percentile(ratings of u, 80)
  recs(q)        user_recs.py:751-774 (similar_user_recs): concatenate the favourites of q's similar users minus
                 q's own favourites, value_counts, first n

The reference identifies anime by `eng_version` names and leaves the order among equal counts to pandas
(value_counts sorts with an unstable sort); here anime are dense indices and ties go to the lower index -- the
documented deviation the CUDA path shares.
"""
import numpy as np


def favourites(indptr, rating, percentile=80.0):
    """bool flag per CSR entry: rating >= np.percentile(that user's ratings, percentile)."""
    fav = np.zeros(len(rating), bool)
    thr = np.zeros(len(indptr) - 1, np.float64)
    for u in range(len(indptr) - 1):
        b, e = int(indptr[u]), int(indptr[u + 1])
        if e > b:
            r = np.asarray(rating[b:e], np.float64)
            thr[u] = np.percentile(r, percentile)
            fav[b:e] = r >= thr[u]
    return fav, thr


def user_recs(indptr, anime_idx, fav, n_anime, query, sim_users, n_recs):
    """(idx, cnt) of the n_recs most common favourites among `sim_users` (row indices, < 0 = padding) that are not
    favourites of `query`; (-1, 0) past the end."""
    cnt = np.zeros(n_anime, np.int64)
    for u in sim_users:
        if u < 0:
            continue
        b, e = int(indptr[u]), int(indptr[u + 1])
        np.add.at(cnt, anime_idx[b:e][fav[b:e]], 1)
    b, e = int(indptr[query]), int(indptr[query + 1])
    cnt[anime_idx[b:e][fav[b:e]]] = 0
    order = np.lexsort((np.arange(n_anime), -cnt))
    order = order[cnt[order] > 0][:n_recs]
    idx = np.full(n_recs, -1, np.int32)
    c = np.zeros(n_recs, np.int32)
    idx[:len(order)] = order
    c[:len(order)] = cnt[order]
    return idx, c
