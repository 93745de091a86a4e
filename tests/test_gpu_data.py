"""data_gpu (device preprocess / index encoding) is bit-equal to data.py, i.e. to the reference's pandas code on
the goldens (tests/golden/preprocess.json, produced by executing preprocess.py), and to the host path on random
frames with duplicates, NaNs and negative zeros."""
import json
import os

import numpy as np
import pytest
import torch

from anime_recommendations_b200 import data

pytestmark = pytest.mark.gpu
GOLD = os.path.join(os.path.dirname(__file__), "golden")

if torch.cuda.is_available():
    from anime_recommendations_b200 import data_gpu


def _raw():
    g = json.load(open(os.path.join(GOLD, "preprocess.json")))
    raw = {k: np.array([np.nan if x is None else x for x in v], dtype=np.float64) for k, v in g["raw"].items()}
    return g, raw


@pytest.mark.parametrize("case", ["plain", "strict"])
def test_preprocess_matches_reference_pandas(case):
    g, raw = _raw()
    c = g["cases"][case]
    out, idx = data_gpu.preprocess_columns(raw, c["args"]["num_reviews"], c["args"]["drop_unwatched"], c["args"]["drop_plan"])
    assert idx.cpu().tolist() == c["index"]
    for col in data.RAW_COLUMNS:
        np.testing.assert_array_equal(out[col].cpu().numpy().astype(np.float64), np.asarray(c["out"][col], np.float64))


@pytest.mark.parametrize("seed", [0, 1])
def test_random_frames_match_the_host_path(seed):
    rng = np.random.RandomState(seed)
    n = 200_000
    raw = dict(user_id=rng.randint(0, 3000, n).astype(np.float64), anime_id=rng.randint(0, 500, n).astype(np.float64),
               rating=rng.randint(0, 11, n).astype(np.float64), watching_status=rng.randint(1, 7, n).astype(np.float64),
               watched_episodes=rng.randint(0, 3, n).astype(np.float64))
    raw["rating"][rng.choice(n, 500, replace=False)] = np.nan
    raw["watched_episodes"][rng.choice(n, 300, replace=False)] = -0.0       # == 0.0 for drop_duplicates and the filter
    dup = rng.choice(n, 5000, replace=False)
    for k in raw:
        raw[k] = np.concatenate([raw[k], raw[k][dup]])
    for kw in (dict(num_reviews=40), dict(num_reviews=25, drop_unwatched=True, drop_plan=True)):
        want, widx = data.preprocess_columns(raw, **kw)
        got, gidx = data_gpu.preprocess_columns(raw, **kw)
        np.testing.assert_array_equal(gidx.cpu().numpy(), widx)
        for col in data.RAW_COLUMNS:
            np.testing.assert_array_equal(got[col].cpu().numpy(), want[col])
    enc_h = data.encode_ratings(want["user_id"].astype(np.int64), want["anime_id"].astype(np.int64), want["rating"])
    enc_g = data_gpu.encode_ratings(got["user_id"].to(torch.int64), got["anime_id"].to(torch.int64), got["rating"])
    for name in ("user", "anime", "rating", "user_ids", "anime_ids"):
        np.testing.assert_array_equal(getattr(enc_g, name), getattr(enc_h, name))


def test_hash_collision_path_is_exact():
    """Force the fallback (every row gets the same hash) and check the exact lexicographic dedup."""
    rng = np.random.RandomState(3)
    mat = torch.from_numpy(rng.randint(0, 3, (5000, 5)).astype(np.float64)).cuda()
    want = data_gpu.first_of_identical_rows(mat).cpu().numpy()
    orig = data_gpu._row_keys
    data_gpu._row_keys = lambda m: (torch.zeros(m.shape[0], dtype=torch.int64, device=m.device), orig(m)[1])
    try:
        got = data_gpu.first_of_identical_rows(mat).cpu().numpy()
    finally:
        data_gpu._row_keys = orig
    np.testing.assert_array_equal(got, want)
    _, first = np.unique(mat.cpu().numpy(), axis=0, return_index=True)
    ref = np.zeros(5000, bool)
    ref[first] = True
    np.testing.assert_array_equal(want, ref)
