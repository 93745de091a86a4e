"""CPU tests pinning the Half-A oracle (oracle/train.py) to what the reference ships."""
import json
import os

import numpy as np
import pytest
import torch

from oracle import train as ot


def test_lrfn_matches_reference_function_and_history_csv(golden_dir):
    g = json.load(open(os.path.join(golden_dir, "lrfn.json")))
    a = g["args"]
    kw = dict(start_lr=float(a["start_lr"]), min_lr=float(a["min_lr"]), max_lr=float(a["max_lr"]),
              rampup_epochs=int(a["rampup_epochs"]), sustain_epochs=int(a["sustain_epochs"]),
              exp_decay=float(a["exp_decay"]))
    ours = [ot.lrfn(e, **kw) for e in range(20)]
    assert ours == g["lrfn"]                       # bit-equal to the reference's lrfn()
    # the shipped history csv: float32 lr logged by Keras for the 15 epochs of the real run
    hist = np.array(g["history_lr"], dtype=np.float32)
    assert len(hist) == g["history_rows"] == 15
    np.testing.assert_array_equal(np.array(ours[:15], dtype=np.float32), hist)
    assert g["history_columns"] == ["loss", "mse", "val_loss", "val_mse", "lr"]
    a2 = g["case2"]["args"]
    kw2 = dict(start_lr=float(a2["start_lr"]), min_lr=float(a2["min_lr"]), max_lr=float(a2["max_lr"]),
               rampup_epochs=int(a2["rampup_epochs"]), sustain_epochs=int(a2["sustain_epochs"]),
               exp_decay=float(a2["exp_decay"]))
    assert [ot.lrfn(e, **kw2) for e in range(10)] == g["case2"]["lrfn"]


def test_pandas_sample_is_randomstate_permutation(golden_dir):
    g = json.load(open(os.path.join(golden_dir, "sample_perm.json")))
    for n, head in g.items():
        perm = np.random.RandomState(42).permutation(int(n))
        assert perm[:len(head)].tolist() == head


def _torch_step_grads(st, iu, ia, t, l2):
    U = torch.tensor(st.U, dtype=torch.float64, requires_grad=True)
    A = torch.tensor(st.A, dtype=torch.float64, requires_grad=True)
    head = torch.tensor(st.head, dtype=torch.float64, requires_grad=True)
    u, a = U[torch.as_tensor(iu)], A[torch.as_tensor(ia)]
    uh = u * torch.rsqrt(torch.clamp((u * u).sum(1, keepdim=True), min=1e-12))
    ah = a * torch.rsqrt(torch.clamp((a * a).sum(1, keepdim=True), min=1e-12))
    c = (uh * ah).sum(1)
    z = head[0] * c + head[1]
    mu = z.mean()
    var = ((z - mu) ** 2).mean()
    y = head[2] * (z - mu) * torch.rsqrt(var + 1e-3) + head[3]
    tt = torch.tensor(t, dtype=torch.float64)
    bce = torch.nn.functional.binary_cross_entropy_with_logits(y, tt)
    loss = bce + l2 * ((U * U).sum() + (A * A).sum())
    loss.backward()
    return float(bce.detach()), float(loss.detach()), U.grad.numpy(), A.grad.numpy(), head.grad.numpy()


def test_manual_backward_matches_autograd():
    rng = np.random.RandomState(0)
    st = ot.init_state(50, 30, 16, seed=1, w=1.3, dtype=np.float64)
    st.head[:] = [1.3, 0.2, 0.9, -0.1]
    B = 200                                            # many duplicate rows
    iu, ia = rng.randint(0, 50, B), rng.randint(0, 30, B)
    t = rng.randint(0, 11, B) / 10.0
    l2 = 1e-4
    bce, loss, gU, gA, gh = _torch_step_grads(st, iu, ia, t, l2)
    fw = ot.forward(st, iu, ia, training=True)
    hb = ot.head_backward(fw, t, st.head, B)
    dc = hb["dc"]
    du = (fw["ru"] * dc)[:, None] * (fw["ah"] - fw["c"][:, None] * fw["uh"])
    da = (fw["ra"] * dc)[:, None] * (fw["uh"] - fw["c"][:, None] * fw["ah"])
    mU = 2 * l2 * st.U
    mA = 2 * l2 * st.A
    np.add.at(mU, iu, du)
    np.add.at(mA, ia, da)
    np.testing.assert_allclose(mU, gU, rtol=1e-9, atol=1e-12)
    np.testing.assert_allclose(mA, gA, rtol=1e-9, atol=1e-12)
    np.testing.assert_allclose(hb["ghead"], gh, rtol=1e-8, atol=1e-12)
    m = ot.train_step(st.clone(), iu, ia, t, 1e-3, l2)
    assert abs(m["bce"] - bce) < 1e-12 and abs(m["loss"] - loss) < 1e-12


def test_train_step_matches_torch_adam_fp32():
    """Whole step (incl. Adam) vs torch.optim.Adam semantics rewritten with Keras' epsilon placement."""
    rng = np.random.RandomState(2)
    st = ot.init_state(40, 25, 8, seed=3, w=-0.8)
    ref = st.clone(np.float64)
    for step in range(5):
        iu, ia = rng.randint(0, 40, 64), rng.randint(0, 25, 64)
        t = rng.randint(0, 11, 64) / 10.0
        m32 = ot.train_step(st, iu, ia, t, 1e-3)
        m64 = ot.train_step(ref, iu, ia, t, 1e-3)
        assert abs(m32["loss"] - m64["loss"]) < 2e-6
    np.testing.assert_allclose(st.U, ref.U, rtol=2e-5, atol=2e-7)
    np.testing.assert_allclose(st.A, ref.A, rtol=2e-5, atol=2e-7)
    # Dense bias: d(loss)/db is identically 0 (BatchNorm removes it); in float32 it is rounding
    # noise that Adam normalises, so `b` random-walks by ~alpha per step.  It cannot affect any
    # output (z - mean(z) cancels it) and is compared loosely everywhere.
    np.testing.assert_allclose(np.delete(st.head, 1), np.delete(ref.head, 1), rtol=2e-5, atol=2e-7)
    assert abs(st.head[1] - ref.head[1]) < 5 * 1e-3 * 5
    assert st.iterations == 5


def test_deferred_replay_equals_dense_adam():
    """replaying the missed pure-L2 Adam steps of an untouched row reproduces dense Adam."""
    f = np.float32
    rng = np.random.RandomState(4)
    l2 = 1e-4
    w = rng.uniform(-0.05, 0.05, 16).astype(f)
    m = (rng.standard_normal(16) * 1e-5).astype(f)
    v = (rng.uniform(0, 1, 16) * 1e-9).astype(f)
    wd, md, vd = w.copy(), m.copy(), v.copy()
    lrs = [1e-5] * 20 + [1.8e-5] * 20
    for i, lr in enumerate(lrs):                      # dense: every step
        ot._adam_apply(wd, md, vd, f(2 * l2) * wd, ot.adam_alpha(lr, 100 + i + 1), f)
    wr, mr, vr = w.copy(), m.copy(), v.copy()
    for i, lr in enumerate(lrs):                      # replay: all at once, element-wise scalars
        for j in range(16):
            g = f(2 * l2) * wr[j]
            mr[j] = mr[j] + (g - mr[j]) * f(1 - ot.BETA1)
            vr[j] = vr[j] + (g * g - vr[j]) * f(1 - ot.BETA2)
            wr[j] = wr[j] - (mr[j] * ot.adam_alpha(lr, 100 + i + 1)) / (np.sqrt(vr[j]) + f(ot.ADAM_EPS))
    np.testing.assert_array_equal(wd, wr)
    np.testing.assert_array_equal(md, mr)
    np.testing.assert_array_equal(vd, vr)


def test_fit_history_and_early_stopping_shape():
    rng = np.random.RandomState(5)
    n = 600
    iu, ia = rng.randint(0, 30, n), rng.randint(0, 20, n)
    y = rng.randint(0, 11, n) / 10.0
    (tu, ta, ty), (vu, va, vy) = ot.split(iu, ia, y, 100)
    assert len(tu) == 500 and len(vu) == 100 and vu[0] == iu[500]
    st = ot.init_state(30, 20, 8, seed=6, w=1.0)
    hist, best = ot.fit(st, [tu, ta], ty, 128, 4, ([vu, va], vy))
    assert list(hist) == ["loss", "mse", "val_loss", "val_mse", "lr"]
    assert len(hist["loss"]) == 4
    assert hist["lr"][:3] == [float(np.float32(ot.lrfn(e))) for e in range(3)]
    assert st.iterations == 4 * 4                     # ceil(500/128) steps per epoch
    assert set(best) == {"U", "A", "head", "mov_mean", "mov_var"}


def test_predict_uses_moving_statistics():
    st = ot.init_state(10, 10, 8, seed=7, w=1.0)
    p0 = ot.predict(st, [0, 1, 2], [3, 4, 5])
    assert p0.shape == (3, 1) and p0.dtype == np.float32
    st.mov_mean, st.mov_var = np.float32(0.5), np.float32(4.0)
    p1 = ot.predict(st, [0, 1, 2], [3, 4, 5])
    assert not np.allclose(p0, p1)


def test_torch_cpu_port_matches_numpy_oracle():
    from oracle import train_torch as tt
    rng = np.random.RandomState(8)
    st = ot.init_state(60, 40, 16, seed=9, w=1.2)
    ts = tt.TorchState(st)
    for _ in range(4):
        iu, ia = rng.randint(0, 60, 128), rng.randint(0, 40, 128)
        t = (rng.randint(0, 11, 128) / 10.0).astype(np.float32)
        m0 = ot.train_step(st, iu, ia, t, 2e-3)
        m1 = tt.train_step(ts, iu, ia, t, 2e-3)
        assert abs(m0["loss"] - m1["loss"]) < 2e-6 and abs(m0["mse"] - m1["mse"]) < 2e-6
    np.testing.assert_allclose(ts.U.numpy(), st.U, rtol=2e-5, atol=2e-7)
    np.testing.assert_allclose(ts.A.numpy(), st.A, rtol=2e-5, atol=2e-7)


def test_oracle_matches_tensorflow_dump(golden_dir):
    """Half A against TensorFlow itself: consumes tests/golden/tf_train.npz when someone has produced it with
    tests/golden/make_tf_golden.py on a machine that has TensorFlow 2.12 (not installable in the build image).
    Until then Half A's parity is UNPINNED: the oracle encodes eight Keras-2.12 assumptions (oracle/train.py) that
    only the lrfn KAT and the torch-autograd cross-check constrain."""
    import os
    path = os.path.join(golden_dir, "tf_train.npz")
    if not os.path.exists(path):
        pytest.skip("parity unpinned: no TensorFlow dump (run tests/golden/make_tf_golden.py where TF 2.12 exists)")
    z = np.load(path)
    st = ot.State(U=z["U0"].copy(), A=z["A0"].copy(), head=z["head0"].copy())
    lr, l2 = float(z["lr"]), float(z["l2"])
    for s in range(z["iu"].shape[0]):
        out = ot.train_step(st, z["iu"][s], z["ia"][s], z["y"][s], lr, l2)
        assert abs(out["loss"] - z["loss"][s]) <= 2e-6 + 2e-6 * abs(z["loss"][s])
        assert abs(out["mse"] - z["mse"][s]) <= 2e-6
    np.testing.assert_allclose(st.U, z["U"], rtol=2e-5, atol=2e-7)
    np.testing.assert_allclose(st.A, z["A"], rtol=2e-5, atol=2e-7)
    np.testing.assert_allclose(np.delete(st.head, 1), np.delete(z["head"], 1), rtol=1e-4, atol=1e-7)
    np.testing.assert_allclose([st.mov_mean, st.mov_var], z["moving"], rtol=1e-5, atol=1e-6)
    np.testing.assert_allclose(ot.predict(st, z["pred_u"], z["pred_a"]).reshape(-1), z["pred"].reshape(-1), atol=2e-5)
    np.testing.assert_allclose(st.mU, z["slot__Adam__m__user_embedding__embeddings:0"], rtol=5e-5, atol=1e-7)
    np.testing.assert_allclose(st.vA, z["slot__Adam__v__anime_embedding__embeddings:0"], rtol=1e-4, atol=1e-11)
