"""GPU edge cases the reference's domain offers: ragged and tiny inputs, maximum sizes, empty
candidate sets, zero-norm rows, duplicates -- CUDA path vs oracle, through the public Python surface."""
import numpy as np
import pytest
import torch

from oracle import similarity as osim
from oracle import train as ot

pytestmark = pytest.mark.gpu

if torch.cuda.is_available():
    import anime_recommendations_b200 as ar
    from anime_recommendations_b200 import _capi, similarity as sim
    from gpu_util import assert_topk_close


def _fit_both(n_users, n_anime, dim, n, B, seed, mode="replay", epochs=1):
    rng = np.random.RandomState(seed)
    iu, ia = rng.randint(0, n_users, n).astype(np.int32), rng.randint(0, n_anime, n).astype(np.int32)
    y = (rng.randint(0, 11, n) / 10.0).astype(np.float32)
    st = ot.init_state(n_users, n_anime, dim, seed=seed, w=0.7)
    m = ar.EmbeddingDotModel(n_users, n_anime, dim, seed=0, adam_mode=mode, dense_kernel=1.0)
    m.set_weights([st.U, st.A, st.head[0:1], st.head[1:2], st.head[2:3], st.head[3:4], np.zeros(1), np.ones(1)])
    m.lr = 1e-3
    m.fit([iu, ia], y, batch_size=B, epochs=epochs, shuffle=False)
    for _ in range(epochs):
        for s in range(0, n, B):
            ot.train_step(st, iu[s:s + B], ia[s:s + B], y[s:s + B], 1e-3)
    w = m.get_weights()
    np.testing.assert_allclose(w[0], st.U, rtol=2e-5, atol=2e-7)
    np.testing.assert_allclose(w[1], st.A, rtol=2e-5, atol=2e-7)
    return m, st


@pytest.mark.parametrize("dim", [4, 12, 512])
def test_training_at_extreme_embedding_sizes(dim):
    _fit_both(90, 40, dim, 700, 256, seed=dim)


def test_training_single_sample_and_fewer_samples_than_batch():
    _fit_both(50, 20, 32, 1, 1000, seed=1)          # one sample: BatchNorm variance 0 -> eps path
    _fit_both(50, 20, 32, 37, 1000, seed=2)         # a single, ragged step


def test_training_at_the_maximum_batch_with_one_hot_row():
    """AR_MAX_BATCH samples in one step, 90 % of them on ONE anime row (the heavy-row CTA path)."""
    B = _capi.AR_MAX_BATCH
    rng = np.random.RandomState(3)
    n_users, n_anime, dim = 3000, 50, 64
    iu = rng.randint(0, n_users, 2 * B).astype(np.int32)
    ia = np.where(rng.rand(2 * B) < 0.9, 7, rng.randint(0, n_anime, 2 * B)).astype(np.int32)
    y = (rng.randint(0, 11, 2 * B) / 10.0).astype(np.float32)
    st = ot.init_state(n_users, n_anime, dim, seed=9, w=1.2)
    m = ar.EmbeddingDotModel(n_users, n_anime, dim, seed=0, dense_kernel=1.0)
    m.set_weights([st.U, st.A, st.head[0:1], st.head[1:2], st.head[2:3], st.head[3:4], np.zeros(1), np.ones(1)])
    m.lr = 5e-4
    m.fit([iu, ia], y, batch_size=B, epochs=1, shuffle=False)
    for s in range(0, 2 * B, B):
        ot.train_step(st, iu[s:s + B], ia[s:s + B], y[s:s + B], 5e-4)
    w = m.get_weights()
    np.testing.assert_allclose(w[1], st.A, rtol=5e-5, atol=5e-7)
    np.testing.assert_allclose(w[0], st.U, rtol=2e-5, atol=2e-7)
    with pytest.raises(ValueError):
        m.fit([iu, ia], y, batch_size=B + 1, epochs=1)


def test_query_topk_edges():
    rng = np.random.RandomState(4)
    W = rng.standard_normal((300, 128)).astype(np.float32)
    W[17] = 0.0                                       # zero-norm row: NaN score in the reference, never ranked here
    W[40] = W[3]                                      # exact duplicate of the query: score 1.0, tie broken by index
    idx, sc = sim.cosine_topk_query(W, 3, 32, exclude=3)
    oi, os_ = osim.rank_desc(osim.query_scores(osim.get_weights(W), 3), 32, exclude=3)
    assert idx.tolist() == oi.tolist() and idx[0] == 40 and 17 not in idx
    np.testing.assert_allclose(sc, os_, atol=3e-6)
    # empty candidate set, k larger than the set, single-row table
    idx, sc = sim.cosine_topk_query(W, 3, 5, mask=np.zeros(300, bool))
    assert idx.size == 0 and sc.size == 0
    m = np.zeros(300, bool)
    m[[5, 9]] = True
    idx, _ = sim.cosine_topk_query(W, 3, 10, mask=m)
    assert sorted(idx.tolist()) == [5, 9]
    idx, sc = sim.cosine_topk_query(W[:1], 0, 3)
    assert idx.tolist() == [0] and abs(sc[0] - 1.0) < 1e-6
    with pytest.raises(KeyError):
        sim.cosine_topk_query(W, 300, 3)
    with pytest.raises(ValueError):
        sim.cosine_topk_query(W, 0, 0)
    idx, sc = sim.cosine_topk_query(W, 0, 33)             # above the kernel's 32 per pass: several passes, same answer
    oi, os_ = osim.rank_desc(osim.query_scores(osim.get_weights(W), 0), 33)
    assert idx.tolist() == oi.tolist()


def test_allpairs_tiny_and_ragged_tables():
    rng = np.random.RandomState(6)
    for n in (2, 11, 129, 257):                       # fewer rows than k, one past a tile, one past a query tile
        W = rng.standard_normal((n, 128)).astype(np.float32)
        gi, gs = sim.allpairs_topk(W, k=10)
        gi, gs = gi.cpu().numpy(), gs.cpu().numpy()
        oi, os_ = osim.allpairs_topk(W, 10)
        Wn = osim.get_weights(W)
        for r in range(n):
            kk = min(10, n - 1)
            assert (gi[r, kk:] == -1).all()
            assert_topk_close(gi[r, :kk], gs[r, :kk], oi[r, :kk], os_[r, :kk], Wn @ Wn[r])


def test_score_topk_user_who_watched_everything_gets_nothing():
    st = ot.init_state(20, 300, 128, seed=1, w=1.0)
    rng = np.random.RandomState(2)
    st.U[:], st.A[:] = rng.standard_normal(st.U.shape), rng.standard_normal(st.A.shape)
    m = ar.EmbeddingDotModel(20, 300, 128, seed=0, dense_kernel=1.0)
    m.set_weights([st.U, st.A, st.head[0:1], st.head[1:2], st.head[2:3], st.head[3:4], np.zeros(1), np.ones(1)])
    indptr = np.array([0, 300, 300])                  # user 4 watched all 300 anime, user 9 none
    widx = np.arange(300, dtype=np.int32)
    gi, gp = sim.score_topk(m, [4, 9], indptr, widx, 5)
    assert (gi[0] == -1).all() and np.isneginf(gp[0]).all()
    oi, op = osim.score_topk(st, [4, 9], indptr, widx, 5)
    assert gi[1].tolist() == oi[1].tolist()
    np.testing.assert_allclose(gp[1], op[1], atol=2e-6)
