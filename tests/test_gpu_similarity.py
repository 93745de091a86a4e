"""GPU parity: cosine top-k (through the C-ABI) against oracle/similarity.py and the golden outputs
of the reference's own find_similar_users / anime_recs.

Tolerance: similarity scores |d| <= 3e-6 (fp32 dot of unit vectors, different summation order than
BLAS); index lists bit-exact except where the oracle's own scores differ by <= that tolerance."""
import os

import numpy as np
import pytest
import torch

from oracle import similarity as osim

pytestmark = pytest.mark.gpu

if torch.cuda.is_available():
    import anime_recommendations_b200 as ar
    from anime_recommendations_b200 import similarity as sim
    from gpu_util import DEV, dev, assert_topk_close, check, lib, ptr, stream_ptr


@pytest.fixture(scope="module")
def world(golden_dir):
    return np.load(os.path.join(golden_dir, "similarity_world.npz"))


def test_rownorm_matches_reference_get_weights(world):
    for key, gold in (("anime_table", "anime_weights_norm"), ("user_table", "user_weights_norm")):
        out = sim.normalize_rows(world[key]).cpu().numpy()
        np.testing.assert_allclose(out, world[gold], rtol=3e-7, atol=0)


def test_similar_users_golden(world):
    uid = world["user_ids"].tolist()
    for q, ids, sims in zip(world["su_query"], world["su_ids"], world["su_sims"]):
        idx, sc = sim.find_similar_users(world["user_table"], uid.index(int(q)), 5)
        assert [uid[i] for i in idx] == ids.tolist()
        np.testing.assert_allclose(sc, sims, rtol=0, atol=3e-6)


def test_similar_anime_golden(world):
    names = ["Anime %d" % i for i in range(len(world["anime_ids"]))]
    tmask = np.isin(world["anime_type"], ["TV", "Movie"])
    idx, sc = sim.similar_anime(world["anime_table"], 4, 10, mask=tmask)
    assert [names[i] for i in idx] == world["sa1_names"].tolist()
    np.testing.assert_allclose(sc, world["sa1_sims"], rtol=0, atol=3e-6)
    idx, sc = sim.similar_anime(world["anime_table"], 17, 7)
    assert [names[i] for i in idx] == world["sa2_names"].tolist()
    g = np.char.replace(np.char.lower(world["anime_genres"]), " ", "")
    mask = np.isin(world["anime_type"], ["TV", "Special", "ONA"]) & (
        (np.char.find(g, "comedy") >= 0) | (np.char.find(g, "vampire") >= 0))
    idx, sc = sim.similar_anime(world["anime_table"], 9, 6, mask=mask)
    assert [names[i] for i in idx] == world["sa3_names"].tolist()
    np.testing.assert_allclose(sc, world["sa3_sims"], rtol=0, atol=3e-6)


@pytest.mark.parametrize("n,dim,k", [(50000, 128, 10), (17560, 128, 32), (3000, 16, 7), (2000, 100, 11),
                                     (4000, 256, 10), (1500, 512, 5), (33, 64, 32), (5, 8, 10),
                                     (3000, 128, 100), (70, 32, 90)])          # k > 32: several passes of the kernel
def test_query_topk_random(n, dim, k):
    rng = np.random.RandomState(n + dim)
    W = rng.standard_normal((n, dim)).astype(np.float32)
    Wd = sim.as_table(W)
    Wn = osim.get_weights(W)
    for q in (0, n // 2, n - 1):
        full = osim.query_scores(Wn, q)
        mask = rng.rand(n) < 0.6
        for kw in (dict(), dict(exclude=q), dict(mask=mask, exclude=q)):
            oi, os_ = osim.rank_desc(full, k, mask=kw.get("mask"), exclude=kw.get("exclude"))
            gi, gs = sim.cosine_topk_query(Wd, q, k, **kw)
            assert_topk_close(gi, gs, oi, os_, full)


def test_query_topk_ties_nan_rows_and_duplicates():
    rng = np.random.RandomState(9)
    W = rng.standard_normal((400, 32)).astype(np.float32)
    W[10] = 0.0                                   # zero-norm row: NaN in the reference, never ranked here
    W[20] = W[7]                                  # exact duplicates of the query: ties broken by lower index
    W[30] = W[7] * 3.0
    gi, gs = sim.cosine_topk_query(W, 7, 5)
    assert gi[:3].tolist() == [7, 20, 30] and 10 not in gi
    gi, gs = sim.find_similar_users(W, 7, 4)      # top-5, query dropped -> 4 rows, duplicates kept
    assert gi[:2].tolist() == [20, 30] and len(gi) == 4
    gi, gs = sim.cosine_topk_query(W, 7, 8, mask=np.zeros(400, bool))
    assert len(gi) == 0
    with pytest.raises(KeyError):
        sim.cosine_topk_query(W, 400, 5)


def test_query_topk_full_user_table_size():
    """cfg3-sized single query: 350 000 x 128."""
    rng = np.random.RandomState(7)
    W = rng.standard_normal((350000, 128)).astype(np.float32)
    Wn = osim.get_weights(W)
    Wd = sim.as_table(W)
    for q in (5, 349999):
        full = osim.query_scores(Wn, q)
        oi, os_ = osim.rank_desc(full, 11)
        gi, gs = sim.cosine_topk_query(Wd, q, 11)
        assert_topk_close(gi, gs, oi, os_, full)


def test_merge_and_rerank():
    rng = np.random.RandomState(3)
    nl, nq, kin, k = 5, 64, 12, 10
    score = rng.standard_normal((nl, nq, kin)).astype(np.float32)
    idx = rng.permutation(nl * nq * kin).reshape(nl, nq, kin).astype(np.int32)
    idx[0, :, -1] = -1                                        # empty slots are skipped
    for sorted_lists in (False, True):
        if sorted_lists:                                      # best-first lists: the bound-pruned merge
            score = -np.sort(-score, axis=2)
            idx[0, :, -1], score[0, :, -1] = -1, -np.inf
        oi, os_ = sim.topk_merge(dev(idx), dev(score), k, lists_sorted=sorted_lists)
        oi, os_ = oi.cpu().numpy(), os_.cpu().numpy()
        for qy in range(nq):
            s = np.where(idx[:, qy].ravel() >= 0, score[:, qy].ravel(), -np.inf)
            ri, rs = osim.rank_desc(s, k)
            np.testing.assert_array_equal(oi[qy], idx[:, qy].ravel()[ri])
            np.testing.assert_array_equal(os_[qy], rs)
    # rerank: candidates = a superset of the true top-k -> exact fp32 top-k back
    W = rng.standard_normal((3000, 128)).astype(np.float32)
    Wn = osim.get_weights(W)
    ti, ts = osim.allpairs_topk_fast(W, k, q0=100, q1=164)
    cand = np.concatenate([ti[:, ::-1], rng.randint(0, 3000, (64, 20)).astype(np.int32)], axis=1)
    cand[:, 15] = -1
    cand[np.arange(64), 12] = np.arange(100, 164)             # the query itself sneaks in: caller filters
    Wd = sim.as_table(W)
    gi, gs, _ = sim.rerank(Wd, 100, 64, Wd, dev(cand), k + 1)
    gi, gs = gi.cpu().numpy(), gs.cpu().numpy()
    for r in range(64):
        keep = gi[r] != 100 + r
        full = Wn @ Wn[100 + r]
        assert_topk_close(gi[r][keep][:k], gs[r][keep][:k], ti[r], ts[r], full)
