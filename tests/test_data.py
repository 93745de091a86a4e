"""Host data plumbing vs goldens produced by executing the reference's own pandas code
(tests/golden/make_golden.py: preprocess.py:13-40,108-117; neural_network.py:43-60)."""
import json
import os

import numpy as np
import pytest

from anime_recommendations_b200 import data

GOLD = os.path.join(os.path.dirname(__file__), "golden")


def _raw():
    g = json.load(open(os.path.join(GOLD, "preprocess.json")))
    raw = {k: np.array([np.nan if x is None else x for x in v], dtype=np.float64) for k, v in g["raw"].items()}
    return g, raw


@pytest.mark.parametrize("case", ["plain", "strict"])
def test_preprocess_matches_reference_pandas(case):
    g, raw = _raw()
    c = g["cases"][case]
    out, idx = data.preprocess_columns(raw, c["args"]["num_reviews"], c["args"]["drop_unwatched"], c["args"]["drop_plan"])
    assert idx.tolist() == c["index"]
    for col in data.RAW_COLUMNS:
        np.testing.assert_array_equal(out[col], np.asarray(c["out"][col], np.float64))
    assert out["rating"].dtype == np.float64 and out["rating"].min() == 0.0 and out["rating"].max() == 1.0


def test_sample_permutation_is_pandas_sample():
    g = json.load(open(os.path.join(GOLD, "sample_perm.json")))
    for n, head in g.items():
        assert data.sample_permutation(int(n))[:64].tolist() == head


def test_first_appearance_vocabulary_and_split():
    ids = np.array([50, 7, 50, 9, 7, 3, 9, 50])
    codes, uniq = data.first_appearance_codes(ids)
    assert uniq.tolist() == [50, 7, 9, 3] and codes.tolist() == [0, 1, 0, 2, 1, 3, 2, 0]
    rng = np.random.RandomState(1)
    u, a = rng.randint(100, 140, 500), rng.randint(1000, 1060, 500)
    y = rng.randint(0, 11, 500) / 10.0
    enc = data.encode_ratings(u, a, y)
    perm = data.sample_permutation(500)
    assert enc.user_ids[enc.user].tolist() == u[perm].tolist()            # decode(encode) == shuffled ids
    assert enc.anime_ids[enc.anime].tolist() == a[perm].tolist()
    np.testing.assert_array_equal(enc.rating, y[perm])
    # vocabulary order = first appearance in the UNSHUFFLED frame
    assert enc.user_ids.tolist() == list(dict.fromkeys(u.tolist()))
    (xtr, ytr), (xte, yte) = data.train_test_split_tail(enc, 100)
    assert len(ytr) == 400 and len(yte) == 100 and xte[0].tolist() == enc.user[400:].tolist()
    with pytest.raises(ValueError):
        data.train_test_split_tail(enc, 500)
