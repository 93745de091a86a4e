"""Index encoding and the preprocess filters on the GPU: the same results as data.py (which is
bit-equal to the reference's pandas on the goldens), computed on device-resident columns so that the 109 M-row
frame of BASELINE cfg2 is filtered and encoded in tens of milliseconds instead of minutes of host hashing.

  preprocess.py:13-40      drop_duplicates (keep first) . dropna . optional filters . >= num_reviews ratings per user
  preprocess.py:108-117    min-max scale of the ratings (float64, the same two IEEE operations)
  neural_network.py:43-60  first-appearance vocabulary of user_id / anime_id, `sample(frac=1, random_state=42)`

The work is sorting and scanning 64-bit keys -- HBM-bound integer work for which the device-wide radix sort behind
torch.sort / torch.unique is the right tool; there is no custom kernel here and none is claimed.  The seeded
permutation itself stays NumPy's Mersenne Twister (a sequential generator; the reference's row order depends on it
bit for bit) and is applied on the device.
"""
from __future__ import annotations

import numpy as np
import torch

from .data import RAW_COLUMNS, EncodedRatings, sample_permutation


def _dev(device):
    if device is not None:
        return torch.device(device)
    if not torch.cuda.is_available():
        from ._capi import AnimerecError
        raise AnimerecError("data_gpu needs a CUDA device; there is no CPU fallback (data.py is the NumPy statement of "
                            "the same rules, used by the tests)")
    return torch.device("cuda", torch.cuda.current_device())


def _i64(x):
    """Python int -> the int64 it wraps to."""
    return ((int(x) + (1 << 63)) % (1 << 64)) - (1 << 63)


def _mix(h):
    """splitmix64 finaliser on int64 tensors (wrapping arithmetic)."""
    h = (h ^ (h >> 30).bitwise_and(0x3FFFFFFFF)) * -4658895280553007687          # 0xBF58476D1CE4E5B9
    h = (h ^ (h >> 27).bitwise_and(0x1FFFFFFFFF)) * -7723592293110705685         # 0x94D049BB133111EB
    return h ^ (h >> 31).bitwise_and(0x1FFFFFFFF)


def _row_keys(mat):
    """(n, 5) float64 -> one int64 hash per row; NaN compares equal to NaN (pandas drop_duplicates semantics)."""
    bits = torch.where(torch.isnan(mat), torch.full_like(mat, float("inf")), mat)
    bits = torch.where(bits == 0, torch.zeros_like(bits), bits).view(torch.int64)      # -0.0 == 0.0
    h = torch.zeros(mat.shape[0], dtype=torch.int64, device=mat.device)
    for c in range(mat.shape[1]):
        h = _mix(h ^ (bits[:, c] + _i64((c + 1) * 0x9E3779B97F4A7C15)))
    return h, bits


def first_of_identical_rows(mat):
    """bool (n,): the row is the first of its group of identical rows (drop_duplicates keep='first')."""
    n = mat.shape[0]
    if n == 0:
        return torch.zeros(0, dtype=torch.bool, device=mat.device)
    h, bits = _row_keys(mat)
    hs, order = torch.sort(h, stable=True)                      # equal rows adjacent, in index order
    b = bits[order]
    same_hash = torch.zeros(n, dtype=torch.bool, device=mat.device)
    same_hash[1:] = hs[1:] == hs[:-1]
    same_row = torch.zeros(n, dtype=torch.bool, device=mat.device)
    same_row[1:] = (b[1:] == b[:-1]).all(dim=1)
    if bool((same_hash & ~same_row).any()):                     # a 64-bit collision: exact lexicographic order instead
        order = torch.arange(n, device=mat.device)
        for c in range(mat.shape[1] - 1, -1, -1):
            order = order[torch.sort(bits[order, c], stable=True)[1]]
        b = bits[order]
        same_row[1:] = (b[1:] == b[:-1]).all(dim=1)
        same_row[0] = False
    keep = torch.zeros(n, dtype=torch.bool, device=mat.device)
    keep[order] = ~same_row
    return keep


def drop_useless(cols, num_reviews, drop_unwatched=False, drop_plan=False, device=None):
    """preprocess.py:13-40 on RAW_COLUMNS (NumPy or torch columns) -> kept row indices, ascending (device int64)."""
    dev = _dev(device)
    mat = torch.stack([torch.as_tensor(np.asarray(cols[c], np.float64) if not torch.is_tensor(cols[c]) else cols[c],
                                       dtype=torch.float64).to(dev) for c in RAW_COLUMNS], dim=1)
    keep = first_of_identical_rows(mat)
    keep &= ~torch.isnan(mat).any(dim=1)
    if drop_unwatched:
        keep &= mat[:, 4] != 0
    if drop_plan:
        keep &= mat[:, 3] != 6
    uid = mat[:, 0].contiguous()
    ids, counts = torch.unique(uid[keep], return_counts=True)                 # value_counts on the survivors
    good = ids[counts >= int(num_reviews)]
    if good.numel():
        pos = torch.searchsorted(good, uid).clamp_(max=good.numel() - 1)
        keep &= good[pos] == uid
    else:
        keep &= False
    return torch.nonzero(keep).reshape(-1)


def scale_ratings(rating):
    """preprocess.py:108-117 in float64 on the device."""
    r = rating.to(torch.float64)
    lo, hi = r.min(), r.max()
    return (r - lo) / (hi - lo)


def preprocess_columns(cols, num_reviews, drop_unwatched=False, drop_plan=False, device=None):
    """drop_useless + scale_ratings -> (dict of surviving device columns, kept indices)."""
    dev = _dev(device)
    idx = drop_useless(cols, num_reviews, drop_unwatched, drop_plan, dev)
    out = {}
    for c in RAW_COLUMNS:
        col = cols[c] if torch.is_tensor(cols[c]) else torch.from_numpy(np.asarray(cols[c]))
        out[c] = col.to(dev)[idx]
    out["rating"] = scale_ratings(out["rating"])
    return out, idx


def first_appearance_codes(ids):
    """(codes int64 (n,), uniques in order of first appearance) -- neural_network.py:43-52 -- on the device."""
    ids = ids.reshape(-1)
    n = ids.numel()
    uniq, inv = torch.unique(ids, return_inverse=True)                      # sorted uniques
    first = torch.full((uniq.numel(),), n, dtype=torch.int64, device=ids.device)
    first.scatter_reduce_(0, inv, torch.arange(n, device=ids.device), reduce="amin")
    order = torch.argsort(first, stable=True)                               # sorted-unique position -> appearance rank
    rank = torch.empty_like(order)
    rank[order] = torch.arange(order.numel(), device=ids.device)
    return rank[inv], uniq[order]


def encode_ratings(user_id, anime_id, rating, random_state=42, device=None, to_host=True):
    """neural_network.py:43-60 on the device.  to_host=True returns the data.EncodedRatings of the host path (NumPy
    arrays); to_host=False keeps (user, anime, rating) as device tensors for EmbeddingDotModel.fit."""
    dev = _dev(device)
    t = lambda x, dt: (x if torch.is_tensor(x) else torch.from_numpy(np.asarray(x))).to(device=dev, dtype=dt)
    u, user_ids = first_appearance_codes(t(user_id, torch.int64))
    a, anime_ids = first_appearance_codes(t(anime_id, torch.int64))
    r = t(rating, torch.float64)
    perm = torch.from_numpy(sample_permutation(u.numel(), random_state)).to(dev)
    u, a, r = u[perm], a[perm], r[perm]
    if to_host:
        return EncodedRatings(u.cpu().numpy(), a.cpu().numpy(), r.cpu().numpy(), user_ids.cpu().numpy(),
                              anime_ids.cpu().numpy())
    return EncodedRatings(u, a, r, user_ids.cpu().numpy(), anime_ids.cpu().numpy())
