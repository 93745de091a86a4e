"""GPU parity of the collaborative aggregation (ar_user_favourites, ar_user_recs) against oracle/user_recs.py and
against the frames of the reference's own similar_user_recs (tests/golden/user_recs.json).  Integer results: exact."""
import numpy as np
import pytest
import torch

from oracle import user_recs as our
from test_oracle_user_recs import check_recs, world

pytestmark = pytest.mark.gpu

if torch.cuda.is_available():
    from anime_recommendations_b200 import user_recs as ur


def _random_csr(rng, n_users, n_anime, lo, hi, grid=True):
    counts = rng.randint(lo, hi, n_users)
    counts[3] = 0                                             # a user without ratings
    counts[5] = min(n_anime, 4000)                            # a heavy user
    u = np.repeat(np.arange(n_users), counts)
    a = np.concatenate([rng.choice(n_anime, c, replace=False) for c in counts]).astype(np.int32)
    r = (rng.randint(0, 11, len(u)) / 10.0) if grid else rng.rand(len(u))
    perm = rng.permutation(len(u))                            # arbitrary file order
    return u[perm], a[perm], r[perm].astype(np.float32)


@pytest.mark.parametrize("grid,pct", [(True, 80.0), (False, 80.0), (True, 75.0), (False, 33.3)])
def test_favourites_match_oracle(grid, pct):
    rng = np.random.RandomState(1)
    n_users, n_anime = 300, 5000
    u, a, r = _random_csr(rng, n_users, n_anime, 1, 400, grid)
    csr = ur.RatingsCSR(u, a, r, n_users, n_anime)
    fav, thr = ur.favourites(csr, pct, return_thresholds=True)
    rs = r[csr.order]
    want, wthr = our.favourites(csr.indptr_host, rs.astype(np.float64), pct)
    np.testing.assert_array_equal(thr.cpu().numpy(), wthr)                  # same double arithmetic as np.percentile
    np.testing.assert_array_equal(fav.cpu().numpy().astype(bool), want)


def test_recs_match_oracle_exactly():
    rng = np.random.RandomState(2)
    n_users, n_anime, k, n_recs = 400, 3001, 10, 20
    u, a, r = _random_csr(rng, n_users, n_anime, 5, 300)
    csr = ur.RatingsCSR(u, a, r, n_users, n_anime)
    fav = ur.favourites(csr)
    q = rng.choice(n_users, 64, replace=False).astype(np.int32)
    sim = rng.randint(0, n_users, (64, k)).astype(np.int32)
    sim[::7, -3:] = -1                                                       # padding
    sim[1] = -1                                                              # no similar users at all
    idx, cnt = ur.similar_user_recs(csr, fav, q, sim, n_recs)
    favh = fav.cpu().numpy().astype(bool)
    ah = csr.anime.cpu().numpy()
    for i in range(64):
        wi, wc = our.user_recs(csr.indptr_host, ah, favh, n_anime, int(q[i]), sim[i].tolist(), n_recs)
        np.testing.assert_array_equal(idx[i].cpu().numpy(), wi)
        np.testing.assert_array_equal(cnt[i].cpu().numpy(), wc)


def test_recs_reproduce_reference_frames():
    g, user_ids, anime_ids, u2i, indptr, a, r = world()
    n_users, n_anime = len(user_ids), len(anime_ids)
    u = np.repeat(np.arange(n_users), np.diff(indptr))
    csr = ur.RatingsCSR(u, a, r, n_users, n_anime)
    fav = ur.favourites(csr)
    favh = fav.cpu().numpy().astype(bool)
    for q, sims, ids, cnts in zip(g["query"], g["sim"], g["rec_anime_id"], g["rec_count"]):
        rows = [u2i[x] for x in sims]
        idx, cnt = ur.similar_user_recs(csr, fav, [u2i[q]], [rows], len(ids))
        allc = np.zeros(n_anime, np.int64)
        for s in rows:
            np.add.at(allc, a[indptr[s]:indptr[s + 1]][favh[indptr[s]:indptr[s + 1]]], 1)
        qi = u2i[q]
        allc[a[indptr[qi]:indptr[qi + 1]][favh[indptr[qi]:indptr[qi + 1]]]] = 0
        check_recs(idx[0].cpu().numpy(), cnt[0].cpu().numpy(), ids, cnts, anime_ids, allc)


def test_user_recs_all_runs_from_the_user_table():
    rng = np.random.RandomState(3)
    n_users, n_anime = 600, 900
    u, a, r = _random_csr(rng, n_users, n_anime, 5, 120)
    csr = ur.RatingsCSR(u, a, r, n_users, n_anime)
    W = torch.randn((n_users, 128), device=csr.device)
    idx, cnt, sim = ur.user_recs_all(W, csr, n_sim_users=8, n_recs=12)
    assert idx.shape == (n_users, 12) and sim.shape == (n_users, 8)
    favh = ur.favourites(csr).cpu().numpy().astype(bool)
    ah = csr.anime.cpu().numpy()
    for q in (0, 17, 599):
        wi, wc = our.user_recs(csr.indptr_host, ah, favh, n_anime, q, sim[q].cpu().tolist(), 12)
        np.testing.assert_array_equal(idx[q].cpu().numpy(), wi)
        np.testing.assert_array_equal(cnt[q].cpu().numpy(), wc)
