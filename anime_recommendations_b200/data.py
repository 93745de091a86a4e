"""Host-side data plumbing either side of the training step.

Mirrors, without pandas in the arithmetic path:
  * preprocess.py:13-40,108-117   drop_duplicates / dropna / optional filters / min-ratings filter / min-max scale
  * neural_network.py:43-60       first-appearance vocabulary, `df.sample(frac=1, random_state=42)`
  * neural_network.py:156-169     last `test_size` shuffled rows = validation

The vocabulary ORDER is contract: every downstream component rebuilds the same maps from the same
preprocessed frame (similar_anime.py:44-52, similar_users.py:42-50, model_recs.py:76-81).
"""
from __future__ import annotations

import numpy as np

RAW_COLUMNS = ("user_id", "anime_id", "rating", "watching_status", "watched_episodes")


def first_appearance_codes(ids):
    """(codes int64 (n,), uniques) with uniques in order of first appearance -- `Series.unique()` +
    `{x: i for i, x in enumerate(unique)}` + `.map` (neural_network.py:43-52)."""
    ids = np.asarray(ids)
    uniq_sorted, first_idx, inv = np.unique(ids, return_index=True, return_inverse=True)
    order = np.argsort(first_idx, kind="stable")          # sorted-unique position -> appearance rank
    rank = np.empty_like(order)
    rank[order] = np.arange(order.size)
    return rank[inv].astype(np.int64), uniq_sorted[order]


def sample_permutation(n, random_state=42):
    """Row order of `df.sample(frac=1, random_state=42)` (neural_network.py:59): pandas draws
    `np.random.RandomState(seed).permutation(n)` (pinned by tests/golden/sample_perm.json)."""
    return np.random.RandomState(random_state).permutation(n)


class EncodedRatings:
    """What `get_df()` returns (neural_network.py:25-63) plus the maps the other components rebuild."""

    def __init__(self, user, anime, rating, user_ids, anime_ids):
        self.user, self.anime, self.rating = user, anime, rating          # shuffled, encoded
        self.user_ids, self.anime_ids = user_ids, anime_ids               # index -> id (first appearance)
        self.n_users, self.n_anime = len(user_ids), len(anime_ids)

    @property
    def user_to_index(self):
        return {int(v): i for i, v in enumerate(self.user_ids)}

    @property
    def anime_to_index(self):
        return {int(v): i for i, v in enumerate(self.anime_ids)}


def encode_ratings(user_id, anime_id, rating, random_state=42):
    """neural_network.py:43-60: encode both id columns by first appearance, then shuffle the rows."""
    u, user_ids = first_appearance_codes(user_id)
    a, anime_ids = first_appearance_codes(anime_id)
    perm = sample_permutation(len(u), random_state)
    return EncodedRatings(u[perm], a[perm], np.asarray(rating, np.float64)[perm], user_ids, anime_ids)


def train_test_split_tail(enc, test_size):
    """neural_network.py:156-169: the LAST `test_size` shuffled rows are the validation set (the second
    `sample(random_state=73)` there reshuffles a frame that is not used again)."""
    n = len(enc.user)
    cut = n - int(test_size)
    if cut <= 0:
        raise ValueError("test_size %d leaves no training rows (n=%d)" % (int(test_size), n))
    tr = ([enc.user[:cut], enc.anime[:cut]], enc.rating[:cut])
    te = ([enc.user[cut:], enc.anime[cut:]], enc.rating[cut:])
    return tr, te


# ------------------------------------------------------------------------------------- preprocess
def drop_useless(cols, num_reviews, drop_unwatched=False, drop_plan=False):
    """preprocess.py:13-40 on a dict of equally long NumPy columns (RAW_COLUMNS).  Returns the kept row
    indices (ascending, i.e. the surviving frame in its original order)."""
    n = len(cols["user_id"])
    mat = np.column_stack([np.asarray(cols[c], np.float64) for c in RAW_COLUMNS])
    # drop_duplicates keeps the FIRST of identical rows; NaN == NaN for this purpose (pandas semantics)
    key = np.where(np.isnan(mat), np.inf, mat)
    _, first = np.unique(key, axis=0, return_index=True)
    keep = np.zeros(n, bool)
    keep[first] = True
    keep &= ~np.isnan(mat).any(axis=1)                                     # dropna
    if drop_unwatched:
        keep &= mat[:, 4] != 0
    if drop_plan:
        keep &= mat[:, 3] != 6
    uid = mat[:, 0]
    ids, counts = np.unique(uid[keep], return_counts=True)                 # value_counts on the survivors
    good = ids[counts >= int(num_reviews)]
    keep &= np.isin(uid, good)
    return np.nonzero(keep)[0]


def scale_ratings(rating):
    """preprocess.py:108-117: (x - min) / (max - min) in float64."""
    r = np.asarray(rating, np.float64)
    lo, hi = r.min(), r.max()
    return (r - lo) / (hi - lo)


def preprocess_columns(cols, num_reviews, drop_unwatched=False, drop_plan=False):
    """drop_useless + scale_ratings -> dict of surviving columns (rating scaled to [0, 1])."""
    idx = drop_useless(cols, num_reviews, drop_unwatched, drop_plan)
    out = {c: np.asarray(cols[c])[idx] for c in RAW_COLUMNS}
    out["rating"] = scale_ratings(out["rating"])
    return out, idx
