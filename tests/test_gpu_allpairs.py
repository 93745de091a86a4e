"""GPU parity: tensor-core (tcgen05) cosine candidates + fp32 re-rank against oracle/similarity.py.

Tolerances: raw bf16-operand scores vs an fp32 matmul of the SAME bf16-rounded operands: |d| <= 2e-6
(only the fp32 accumulation order differs); final all-pairs top-k: fp32 scores |d| <= 3e-6 and index
lists bit-exact except where the oracle's own scores tie within that tolerance."""
import numpy as np
import pytest
import torch

from oracle import similarity as osim
from oracle import train as ot

pytestmark = pytest.mark.gpu

if torch.cuda.is_available():
    import anime_recommendations_b200 as ar
    from anime_recommendations_b200 import similarity as sim
    from gpu_util import DEV, dev, assert_topk_close


def _bf16_round(x):
    return torch.from_numpy(x).to(torch.bfloat16).float().numpy()


@pytest.mark.parametrize("nq,nc,q0,c0", [(300, 500, 0, 0), (256, 128, 0, 0), (700, 3000, 37, 129), (1, 1, 0, 0)])
def test_tcgen05_scores_match_matmul_of_same_bf16_operands(nq, nc, q0, c0):
    rng = np.random.RandomState(nq + nc)
    Q = rng.standard_normal((nq + q0 + 5, 128)).astype(np.float32)
    C = rng.standard_normal((nc + c0 + 3, 128)).astype(np.float32)
    Qn, Cn = sim.normalize_rows_bf16(Q), sim.normalize_rows_bf16(C)
    got, want = Qn.float().cpu().numpy(), _bf16_round(osim.get_weights(Q))
    np.testing.assert_allclose(got, want, rtol=2 ** -7, atol=0)      # a 1-ulp fp32 difference can flip a bf16 rounding
    assert (got != want).mean() < 1e-3
    cl = sim.allpairs_candidates(Qn, q0, nq, Cn, c0, nc, kprime=16, dump=True)
    ref = Qn.float().cpu().numpy()[q0:q0 + nq] @ Cn.float().cpu().numpy()[c0:c0 + nc].T
    d = cl.dump.cpu().numpy()
    np.testing.assert_allclose(d, ref, rtol=0, atol=2e-6)
    # list invariant per (chunk, row): the list is EXACTLY the set of chunk candidates scoring above the final
    # threshold, it holds the 16 best of the chunk, and it never exceeds the capacity
    ci, cs, cc, ct = (x.cpu().numpy() for x in (cl.idx, cl.score, cl.cnt, cl.thr))
    n_chunks, _, cap = ci.shape
    tiles = (nc + 127) // 128
    per = (tiles + n_chunks - 1) // n_chunks
    for ch in range(n_chunks):
        lo, hi = ch * per * 128, min(nc, (ch + 1) * per * 128)
        for r in range(0, nq, max(1, nq // 50)):
            cnt, thr = int(cc[ch, r]), float(ct[ch, r])
            assert min(12, hi - lo) <= cnt <= cap            # a compaction keeps 16 unless scores tie at the cut
            sel = ci[ch, r, :cnt]
            assert (ci[ch, r, cnt:] == -1).all() and (sel >= c0 + lo).all() and (sel < c0 + hi).all()
            np.testing.assert_array_equal(cs[ch, r, :cnt], d[r, sel - c0])
            above = np.nonzero(d[r, lo:hi] > thr)[0] + lo + c0
            np.testing.assert_array_equal(np.sort(sel), above)
            top = np.argsort(-d[r, lo:hi], kind="stable")[:min(cnt, 12)] + lo + c0
            assert set(top) <= set(sel)


def test_candidate_lists_honour_thr_init_self_ids_and_watched():
    rng = np.random.RandomState(5)
    W = rng.standard_normal((1500, 128)).astype(np.float32)
    Wn = sim.normalize_rows_bf16(W)
    rows = torch.tensor([7, 300, 1499, 42], dtype=torch.int64, device=DEV)
    Qb = Wn[rows].contiguous()
    thr0 = torch.tensor([0.15, -1.0, 0.3, 0.2], dtype=torch.float32, device=DEV)
    watched = torch.zeros((4, (1500 + 31) // 32), dtype=torch.int32, device=DEV)
    watched[3, 0] = 0x7fffffff                                        # query 3 has "watched" rows 0..30
    cl = sim.allpairs_candidates(Qb, 0, 4, Wn, 0, 1500, kprime=16, exclude_self=True,
                                 self_ids=rows.to(torch.int32), watched=watched, thr_init=thr0, dump=True, n_chunks=1)
    d = cl.dump.cpu().numpy()
    for i in range(4):
        cnt, thr = int(cl.cnt[0, i]), float(cl.thr[0, i])
        assert thr >= float(thr0[i])
        ok = np.ones(1500, bool)
        ok[int(rows[i])] = False
        if i == 3:
            ok[:31] = False
        want = np.nonzero((d[i] > thr) & ok)[0]
        np.testing.assert_array_equal(np.sort(cl.idx[0, i, :cnt].cpu().numpy()), want)


@pytest.mark.parametrize("n,kprime", [(3000, 16), (18000, 16), (5000, 24)])
def test_allpairs_topk_matches_oracle(n, kprime):
    rng = np.random.RandomState(n)
    W = rng.standard_normal((n, 128)).astype(np.float32)
    W[11] = W[5] * 2.0                                     # exact duplicate direction: score 1.0 tie pair
    st = {}
    gi, gs = sim.allpairs_topk(W, k=10, kprime=kprime, stats=st)
    gi, gs = gi.cpu().numpy(), gs.cpu().numpy()
    oi, os_ = osim.allpairs_topk_fast(W, 10)
    Wn = osim.get_weights(W)
    bad = np.nonzero((gi != oi).any(axis=1))[0]
    np.testing.assert_allclose(gs, os_, rtol=0, atol=3e-6)
    for r in bad:                                          # differences only at fp32 score ties
        assert_topk_close(gi[r], gs[r], oi[r], os_[r], Wn @ Wn[r])
    assert len(bad) <= n // 100
    assert (gi != np.arange(n)[:, None]).all()             # self never recommended
    assert st["uncertified"] <= n // 50


def test_allpairs_query_subrange_and_clustered_rows_fall_back_exactly():
    """Tight clusters make bf16 scores indistinguishable: rows that cannot be certified take the fp32 path."""
    rng = np.random.RandomState(1)
    base = rng.standard_normal((40, 128)).astype(np.float32)
    W = (base[rng.randint(0, 40, 4000)] + 1e-3 * rng.standard_normal((4000, 128))).astype(np.float32)
    st = {}
    gi, gs = sim.allpairs_topk(W, k=10, kprime=16, q0=1000, nq=300, stats=st)
    oi, os_ = osim.allpairs_topk_fast(W, 10, q0=1000, q1=1300)
    assert st["uncertified"] > 0
    Wn = osim.get_weights(W)
    gi, gs = gi.cpu().numpy(), gs.cpu().numpy()
    np.testing.assert_allclose(gs, os_, rtol=0, atol=3e-6)
    for r in np.nonzero((gi != oi).any(axis=1))[0]:
        assert_topk_close(gi[r], gs[r], oi[r], os_[r], Wn @ Wn[1000 + r])


@pytest.mark.parametrize("w", [1.4, -0.9])
def test_score_topk_matches_oracle_model_recs(w):
    rng = np.random.RandomState(3)
    nu, na, k = 500, 2100, 20
    st = ot.init_state(nu, na, 128, seed=4, w=w)
    st.U[:] = rng.standard_normal(st.U.shape).astype(np.float32)
    st.A[:] = rng.standard_normal(st.A.shape).astype(np.float32)
    st.head[:] = [w, 0.05, 0.8, -0.1]
    st.mov_mean, st.mov_var = np.float32(0.02), np.float32(0.03)
    m = ar.EmbeddingDotModel(nu, na, 128, seed=0, dense_kernel=1.0)
    m.set_weights([st.U, st.A, st.head[0:1], st.head[1:2], st.head[2:3], st.head[3:4],
                   np.array([st.mov_mean]), np.array([st.mov_var])])
    users = rng.choice(nu, 300, replace=False)
    counts = rng.randint(400, 1500, len(users))
    indptr = np.r_[0, np.cumsum(counts)]
    widx = np.concatenate([rng.choice(na, c, replace=False) for c in counts]).astype(np.int32)
    cand_mask = rng.rand(na) < 0.7
    gi, gp = sim.score_topk(m, users, indptr, widx, k, cand_mask=cand_mask)
    oi, op = osim.score_topk(st, users, indptr, widx, k, cand_mask=cand_mask)
    np.testing.assert_allclose(gp, op, rtol=0, atol=2e-6)
    for r in np.nonzero((gi != oi).any(axis=1))[0]:
        full = osim.model_scores(st, users[r], np.arange(na))
        assert_topk_close(gi[r], gp[r], oi[r], op[r], full, tol=2e-6)
    for j in range(len(users)):                            # nothing watched or filtered out is recommended
        assert not set(gi[j]) & set(widx[indptr[j]:indptr[j + 1]])
        assert cand_mask[gi[j]].all()


@pytest.mark.parametrize("world", [2, 3])
def test_score_topk_sharded_over_users_equals_unsharded(world):
    """BASELINE cfg4 / SURVEY 8e row 4: the query users are split over ranks, nothing else is shared -- the ranks'
    results laid end to end are the unsharded result (here the ranks run one after the other on this GPU; the
    2-process version is tests/test_gpu_dist.py::test_two_rank_sharded_scoring)."""
    from anime_recommendations_b200 import similarity_dist as sd
    rng = np.random.RandomState(5)
    nu, na, k = 400, 1500, 20
    m = ar.EmbeddingDotModel(nu, na, 128, seed=0, dense_kernel=-0.7)
    users = rng.choice(nu, 257, replace=False)
    counts = rng.randint(100, 600, len(users))
    indptr = np.r_[0, np.cumsum(counts)]
    widx = np.concatenate([rng.choice(na, c, replace=False) for c in counts]).astype(np.int32)
    wi, wp = sim.score_topk(m, users, indptr, widx, k)
    parts = [sd.score_topk_sharded(m, users, indptr, widx, k, r, world) for r in range(world)]
    assert parts[0][0] == 0 and parts[-1][1] == len(users)
    assert all(parts[r][1] == parts[r + 1][0] for r in range(world - 1))
    np.testing.assert_array_equal(np.concatenate([p[2] for p in parts]), wi)
    np.testing.assert_array_equal(np.concatenate([p[3] for p in parts]), wp)


@pytest.mark.parametrize("dim,expect_tensor", [(64, True), (100, True), (256, False)])
def test_allpairs_other_embedding_sizes(dim, expect_tensor):
    """config.yaml:63 makes the embedding size a knob: smaller sizes ride the tensor-core pass zero-padded to 128
    (zero columns change no cosine), larger ones take the exact fp32 kernel for every row in ONE C call."""
    rng = np.random.RandomState(dim)
    n = 1500
    W = rng.standard_normal((n, dim)).astype(np.float32)
    st = {}
    gi, gs = sim.allpairs_topk(W, k=10, stats=st)
    gi, gs = gi.cpu().numpy(), gs.cpu().numpy()
    oi, os_ = osim.allpairs_topk_fast(W, 10)
    np.testing.assert_allclose(gs, os_, rtol=0, atol=3e-6)
    Wn = osim.get_weights(W)
    for r in np.nonzero((gi != oi).any(axis=1))[0]:
        assert_topk_close(gi[r], gs[r], oi[r], os_[r], Wn @ Wn[r])
    assert ("uncertified" in st) == expect_tensor             # the certified tensor-core pipeline ran (or not)


def test_score_topk_small_embedding_uses_the_tensor_path():
    rng = np.random.RandomState(9)
    nu, na, k, D = 300, 1200, 20, 64
    st = ot.init_state(nu, na, D, seed=4, w=1.1)
    st.U[:] = rng.standard_normal(st.U.shape).astype(np.float32)
    st.A[:] = rng.standard_normal(st.A.shape).astype(np.float32)
    st.head[:] = [1.1, 0.05, 0.8, -0.1]
    st.mov_mean, st.mov_var = np.float32(0.02), np.float32(0.03)
    m = ar.EmbeddingDotModel(nu, na, D, seed=0, dense_kernel=1.0)
    m.set_weights([st.U, st.A, st.head[0:1], st.head[1:2], st.head[2:3], st.head[3:4],
                   np.array([st.mov_mean]), np.array([st.mov_var])])
    users = rng.choice(nu, 150, replace=False)
    counts = rng.randint(100, 500, len(users))
    indptr = np.r_[0, np.cumsum(counts)]
    widx = np.concatenate([rng.choice(na, c, replace=False) for c in counts]).astype(np.int32)
    stats = {}
    gi, gp = sim.score_topk(m, users, indptr, widx, k, stats=stats)
    oi, op = osim.score_topk(st, users, indptr, widx, k)
    assert "uncertified" in stats
    np.testing.assert_allclose(gp, op, rtol=0, atol=2e-6)
    for r in np.nonzero((gi != oi).any(axis=1))[0]:
        full = osim.model_scores(st, users[r], np.arange(na))
        assert_topk_close(gi[r], gp[r], oi[r], op[r], full, tol=2e-6)


def test_allpairs_full_user_table_sampled_against_oracle():
    """cfg3 at full size (350 000 x 128): every 5000th row checked against the single-query oracle."""
    rng = np.random.RandomState(7)
    n = 350000
    W = rng.standard_normal((n, 128)).astype(np.float32)
    st = {}
    gi, gs = sim.allpairs_topk(W, k=10, kprime=16, stats=st)
    gi, gs = gi.cpu().numpy(), gs.cpu().numpy()
    assert st["uncertified"] < n // 20 and st["uncertified_after_retry"] < 50
    Wn = osim.get_weights(W)
    for r in range(0, n, 5000):
        full = Wn @ Wn[r]
        oi, os_ = osim.rank_desc(full, 10, exclude=r)
        assert_topk_close(gi[r], gs[r], oi, os_, full)
    assert (gi >= 0).all() and (gi != np.arange(n)[:, None]).all()
    assert np.all(np.diff(gs, axis=1) <= 0)
