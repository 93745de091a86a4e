"""Oracle for Half A: the neural_network.py training step (TEST INFRASTRUCTURE).

Restates reference neural_network/neural_network.py:66-106 (model), :109-125
(lrfn), :156-169 (split), :184-217 (fit + callbacks) in NumPy.  The reference
delegates the arithmetic to TensorFlow 2.12.0 / Keras 2.12 (requirements.txt:22,
neural_network/conda.yml:20-21), which is not vendored and not installable here,
so the following Keras-2.12 semantics are ASSUMPTIONS ("[K2.12]" in SURVEY.md):

 A1  Embedding init U(-0.05, 0.05); Dense(1) kernel he_normal with fan_in=1
     (truncated normal, stddev sqrt(2)/0.87962566), bias 0.
 A2  Dot(normalize=True): x * rsqrt(max(sum(x^2), 1e-12)) on both operands.
 A3  BatchNormalization on a rank-2 input takes the non-fused path: batch mean,
     biased batch variance, eps 1e-3, momentum 0.99, gamma=1, beta=0,
     moving_mean=0, moving_variance=1; the moving variance is updated with the
     biased batch variance; mov -= (mov - batch) * (1 - momentum).
 A4  binary_crossentropy after a sigmoid Activation is evaluated from the cached
     logits: max(y,0) - y*t + log1p(exp(-|y|)), reduced by sum_over_batch_size.
 A5  L2 regulariser lam*sum(W^2) on both tables is part of `loss` and
     `val_loss`; its gradient 2*lam*W is dense, so the table gradients are dense.
 A6  optimizer='Adam' -> keras.optimizers.Adam(lr=1e-3 overwritten per epoch by
     LearningRateScheduler, beta_1=0.9, beta_2=0.999, epsilon=1e-7), t =
     iterations+1, alpha = lr*sqrt(1-b2^t)/(1-b1^t), m += (g-m)(1-b1),
     v += (g^2-v)(1-b2), theta -= alpha*m/(sqrt(v)+eps), applied DENSELY to every
     row of both tables and to the 4 head scalars every step.
 A7  fit(): reshuffle every epoch, last batch partial, epoch `loss`/`mse` are
     sample-weighted running means, validation runs in inference mode (moving
     statistics) after every epoch, History keys loss,mse,val_loss,val_mse,lr.
 A8  EarlyStopping(patience=3, restore_best_weights=True) restores the best
     weights only if it actually stops the run early; ModelCheckpoint keeps the
     weights of the best val_loss epoch.

Pinned pieces: `lrfn` (vs figure_file/anime_nn_history.csv) and the manual
backward (vs torch autograd) -- see tests/test_oracle_train.py.
"""
from __future__ import annotations

import copy
from dataclasses import dataclass, field

import numpy as np

BETA1 = 0.9
BETA2 = 0.999
ADAM_EPS = 1e-7
BN_EPS = 1e-3
BN_MOMENTUM = 0.99
L2NORM_EPS = 1e-12

# head scalar order used everywhere (oracle, C-ABI, saved weights)
HEAD_W, HEAD_B, HEAD_GAMMA, HEAD_BETA = 0, 1, 2, 3


def lrfn(epoch, start_lr=1e-5, min_lr=1e-5, max_lr=5e-5, rampup_epochs=5,
         sustain_epochs=0, exp_decay=0.8):
    """Learning-rate schedule, reference neural_network.py:109-125."""
    if epoch < rampup_epochs:
        return (max_lr - start_lr) / rampup_epochs * epoch + start_lr
    if epoch < rampup_epochs + sustain_epochs:
        return max_lr
    return (max_lr - min_lr) * exp_decay ** (epoch - rampup_epochs - sustain_epochs) + min_lr


def adam_alpha(lr, t, dtype=np.float32):
    """alpha_t of Keras-2.12 Adam (A6); t is 1-based."""
    f = dtype
    b1p = np.power(f(BETA1), f(t))
    b2p = np.power(f(BETA2), f(t))
    return f(f(lr) * np.sqrt(f(1) - b2p) / (f(1) - b1p))


@dataclass
class State:
    U: np.ndarray
    A: np.ndarray
    head: np.ndarray                      # [w, b, gamma, beta]
    mov_mean: float = 0.0
    mov_var: float = 1.0
    iterations: int = 0
    mU: np.ndarray = None
    vU: np.ndarray = None
    mA: np.ndarray = None
    vA: np.ndarray = None
    mh: np.ndarray = None
    vh: np.ndarray = None
    dtype: type = np.float32
    extras: dict = field(default_factory=dict)

    def __post_init__(self):
        f = self.dtype
        self.U = np.ascontiguousarray(self.U, dtype=f)
        self.A = np.ascontiguousarray(self.A, dtype=f)
        self.head = np.asarray(self.head, dtype=f).copy()
        self.mov_mean = f(self.mov_mean)
        self.mov_var = f(self.mov_var)
        for name, ref in (("mU", self.U), ("vU", self.U), ("mA", self.A),
                          ("vA", self.A), ("mh", self.head), ("vh", self.head)):
            if getattr(self, name) is None:
                setattr(self, name, np.zeros_like(ref))
            else:
                setattr(self, name, np.ascontiguousarray(getattr(self, name), dtype=f))

    def weights(self):
        """What Keras `model.get_weights()` holds (no optimizer slots)."""
        return dict(U=self.U.copy(), A=self.A.copy(), head=self.head.copy(),
                    mov_mean=self.mov_mean, mov_var=self.mov_var)

    def set_weights(self, w):
        self.U[...] = w["U"]
        self.A[...] = w["A"]
        self.head[...] = w["head"]
        self.mov_mean = self.dtype(w["mov_mean"])
        self.mov_var = self.dtype(w["mov_var"])

    def clone(self, dtype=None):
        s = copy.deepcopy(self)
        if dtype is not None and dtype is not self.dtype:
            s = State(U=s.U, A=s.A, head=s.head, mov_mean=s.mov_mean, mov_var=s.mov_var,
                      iterations=s.iterations, mU=s.mU, vU=s.vU, mA=s.mA, vA=s.vA,
                      mh=s.mh, vh=s.vh, dtype=dtype)
        return s


def he_normal_scalar(rng):
    """Dense(1) kernel of shape (1,1), he_normal: truncated normal (|x|<=2 sd)."""
    sd = np.sqrt(2.0 / 1.0) / 0.87962566103423978
    while True:
        x = rng.standard_normal()
        if abs(x) <= 2.0:
            return np.float32(x * sd)


def init_state(n_users, n_anime, dim, seed=0, w=None, dtype=np.float32):
    """Fresh model as `neural_network()` builds it (neural_network.py:66-106, A1/A3)."""
    rng = np.random.RandomState(seed)
    U = rng.uniform(-0.05, 0.05, size=(n_users, dim)).astype(np.float32)
    A = rng.uniform(-0.05, 0.05, size=(n_anime, dim)).astype(np.float32)
    if w is None:
        w = he_normal_scalar(rng)
    head = np.array([w, 0.0, 1.0, 0.0], dtype=np.float32)
    return State(U=U, A=A, head=head, dtype=dtype)


def _sigmoid(y):
    out = np.empty_like(y)
    pos = y >= 0
    out[pos] = 1.0 / (1.0 + np.exp(-y[pos]))
    e = np.exp(y[~pos])
    out[~pos] = e / (1.0 + e)
    return out


def forward(st: State, iu, ia, training: bool):
    """Model forward, neural_network.py:73-100 (A2, A3)."""
    f = st.dtype
    u = st.U[iu]
    a = st.A[ia]
    ru = f(1) / np.sqrt(np.maximum(np.sum(u * u, axis=1, dtype=f), f(L2NORM_EPS)))
    ra = f(1) / np.sqrt(np.maximum(np.sum(a * a, axis=1, dtype=f), f(L2NORM_EPS)))
    uh = u * ru[:, None]
    ah = a * ra[:, None]
    c = np.sum(uh * ah, axis=1, dtype=f)
    w, b, gamma, beta = (st.head[i] for i in range(4))
    z = w * c + b
    if training:
        mu = np.mean(z, dtype=f)
        var = np.mean((z - mu) ** 2, dtype=f)
    else:
        mu, var = st.mov_mean, st.mov_var
    inv = f(1) / np.sqrt(var + f(BN_EPS))
    zh = (z - mu) * inv
    y = gamma * zh + beta
    p = _sigmoid(y)
    return dict(u=u, a=a, ru=ru, ra=ra, uh=uh, ah=ah, c=c, z=z, mu=mu, var=var,
                inv=inv, zh=zh, y=y, p=p)


def bce_from_logits(y, t):
    return np.maximum(y, 0) - y * t + np.log1p(np.exp(-np.abs(y)))


def reg_loss(st: State, l2):
    f = st.dtype
    return f(l2) * (np.sum(st.U * st.U, dtype=f) + np.sum(st.A * st.A, dtype=f))


def head_backward(fw, t, head, n):
    """Backward through sigmoid/BCE/BN/Dense down to d(loss)/dc."""
    f = fw["y"].dtype.type
    w, b, gamma, beta = (head[i] for i in range(4))
    dy = (fw["p"] - t) / f(n)
    zh = fw["zh"]
    dgamma = np.sum(dy * zh, dtype=f)
    dbeta = np.sum(dy, dtype=f)
    dzh = gamma * dy
    s1 = np.sum(dzh, dtype=f)
    s2 = np.sum(dzh * zh, dtype=f)
    dz = fw["inv"] / f(n) * (f(n) * dzh - s1 - zh * s2)
    dw = np.sum(dz * fw["c"], dtype=f)
    db = np.sum(dz, dtype=f)
    dc = w * dz
    return dict(dc=dc, ghead=np.array([dw, db, dgamma, dbeta], dtype=f))


def _adam_apply(theta, m, v, g, alpha, f):
    m += (g - m) * f(1 - BETA1)
    v += (g * g - v) * f(1 - BETA2)
    theta -= (m * alpha) / (np.sqrt(v) + f(ADAM_EPS))


def train_step(st: State, iu, ia, t, lr, l2=1e-4):
    """One Keras train_step: forward, loss, dense gradients, dense Adam (A2-A6).

    Mutates `st`; returns the step's metrics (computed with the PRE-update weights,
    as Keras reports them).
    """
    f = st.dtype
    iu = np.asarray(iu, dtype=np.int64)
    ia = np.asarray(ia, dtype=np.int64)
    t = np.asarray(t, dtype=f)
    n = len(iu)
    fw = forward(st, iu, ia, training=True)
    bce = np.mean(bce_from_logits(fw["y"], t), dtype=f)
    mse = np.mean((t - fw["p"]) ** 2, dtype=f)
    reg = reg_loss(st, l2)

    hb = head_backward(fw, t, st.head, n)
    dc = hb["dc"]
    du = (fw["ru"] * dc)[:, None] * (fw["ah"] - fw["c"][:, None] * fw["uh"])
    da = (fw["ra"] * dc)[:, None] * (fw["uh"] - fw["c"][:, None] * fw["ah"])
    gU = f(2 * l2) * st.U
    gA = f(2 * l2) * st.A
    np.add.at(gU, iu, du)
    np.add.at(gA, ia, da)

    step = st.iterations + 1
    alpha = adam_alpha(lr, step, f)
    _adam_apply(st.U, st.mU, st.vU, gU, alpha, f)
    _adam_apply(st.A, st.mA, st.vA, gA, alpha, f)
    _adam_apply(st.head, st.mh, st.vh, hb["ghead"], alpha, f)
    st.mov_mean = f(st.mov_mean - (st.mov_mean - fw["mu"]) * f(1 - BN_MOMENTUM))
    st.mov_var = f(st.mov_var - (st.mov_var - fw["var"]) * f(1 - BN_MOMENTUM))
    st.iterations = step
    return dict(bce=float(bce), reg=float(reg), loss=float(bce + reg), mse=float(mse),
                n=n, mu=float(fw["mu"]), var=float(fw["var"]))


def train_step_touched_only(st: State, iu, ia, t, lr, l2=1e-4):
    """The north-star's literal 'Adam on touched rows only' variant: rows absent from
    the batch are left completely alone (no L2 pull, no moment decay).  This is NOT
    the reference's arithmetic; it is the oracle for the library's
    `touched` mode only."""
    f = st.dtype
    iu = np.asarray(iu, dtype=np.int64)
    ia = np.asarray(ia, dtype=np.int64)
    t = np.asarray(t, dtype=f)
    n = len(iu)
    fw = forward(st, iu, ia, training=True)
    bce = np.mean(bce_from_logits(fw["y"], t), dtype=f)
    mse = np.mean((t - fw["p"]) ** 2, dtype=f)
    hb = head_backward(fw, t, st.head, n)
    dc = hb["dc"]
    du = (fw["ru"] * dc)[:, None] * (fw["ah"] - fw["c"][:, None] * fw["uh"])
    da = (fw["ra"] * dc)[:, None] * (fw["uh"] - fw["c"][:, None] * fw["ah"])
    step = st.iterations + 1
    alpha = adam_alpha(lr, step, f)
    for W, m, v, idx, d in ((st.U, st.mU, st.vU, iu, du), (st.A, st.mA, st.vA, ia, da)):
        rows, inv = np.unique(idx, return_inverse=True)
        g = np.zeros((len(rows), W.shape[1]), dtype=f)
        np.add.at(g, inv, d)
        g += f(2 * l2) * W[rows]
        wr, mr, vr = W[rows], m[rows], v[rows]
        _adam_apply(wr, mr, vr, g, alpha, f)
        W[rows], m[rows], v[rows] = wr, mr, vr
    _adam_apply(st.head, st.mh, st.vh, hb["ghead"], alpha, f)
    st.mov_mean = f(st.mov_mean - (st.mov_mean - fw["mu"]) * f(1 - BN_MOMENTUM))
    st.mov_var = f(st.mov_var - (st.mov_var - fw["var"]) * f(1 - BN_MOMENTUM))
    st.iterations = step
    return dict(bce=float(bce), mse=float(mse), n=n)


def predict(st: State, iu, ia):
    """`model.predict([users, animes])` -> (M,1) float32 (model_recs.py:394), inference BN."""
    fw = forward(st, np.asarray(iu, dtype=np.int64), np.asarray(ia, dtype=np.int64), training=False)
    return fw["p"].reshape(-1, 1)


def evaluate(st: State, iu, ia, t, batch_size, l2=1e-4):
    """Keras validation pass: inference mode, sample-weighted means, reg included (A5, A7)."""
    f = st.dtype
    t = np.asarray(t, dtype=f)
    n = len(iu)
    sb = sm = 0.0
    for s in range(0, n, batch_size):
        fw = forward(st, iu[s:s + batch_size], ia[s:s + batch_size], training=False)
        tt = t[s:s + batch_size]
        sb += float(np.sum(bce_from_logits(fw["y"], tt), dtype=np.float64))
        sm += float(np.sum((tt - fw["p"]) ** 2, dtype=np.float64))
    reg = float(reg_loss(st, l2))
    return dict(val_bce=sb / n, val_loss=sb / n + reg, val_mse=sm / n, reg=reg)


def split(users, animes, ratings, test_size):
    """neural_network.py:156-169: the last `test_size` shuffled rows are validation."""
    k = len(users) - int(test_size)
    return (users[:k], animes[:k], ratings[:k]), (users[k:], animes[k:], ratings[k:])


def fit(st: State, x, y, batch_size, epochs, validation_data, lr_kwargs=None, l2=1e-4,
        shuffle_seed=0, patience=3, step_fn=train_step, on_step=None):
    """`model.fit` + LearningRateScheduler + ModelCheckpoint(best) + EarlyStopping (A7, A8).

    Returns (history dict, best_weights).  The per-epoch shuffle is
    np.random.RandomState(shuffle_seed + epoch).permutation(n) -- Keras' shuffle is
    unseeded, so the library uses the same documented rule to stay comparable.
    """
    lr_kwargs = lr_kwargs or {}
    iu, ia = (np.asarray(a) for a in x)
    y = np.asarray(y, dtype=st.dtype)
    (vu, va), vy = validation_data
    n = len(iu)
    hist = {k: [] for k in ("loss", "mse", "val_loss", "val_mse", "lr")}
    best, best_w, ckpt_w, wait = np.inf, None, None, 0
    for epoch in range(epochs):
        lr = lrfn(epoch, **lr_kwargs)
        perm = np.random.RandomState(shuffle_seed + epoch).permutation(n)
        sl = ss = 0.0
        for s in range(0, n, batch_size):
            b = perm[s:s + batch_size]
            m = step_fn(st, iu[b], ia[b], y[b], lr, l2)
            if on_step is not None:
                on_step(st, m)
            sl += m.get("loss", m["bce"]) * m["n"]
            ss += m["mse"] * m["n"]
        ev = evaluate(st, vu, va, vy, batch_size, l2)
        hist["loss"].append(sl / n)
        hist["mse"].append(ss / n)
        hist["val_loss"].append(ev["val_loss"])
        hist["val_mse"].append(ev["val_mse"])
        hist["lr"].append(float(np.float32(lr)))
        if best_w is None:
            best_w = st.weights()
        wait += 1
        if ev["val_loss"] < best:
            best, best_w, ckpt_w, wait = ev["val_loss"], st.weights(), st.weights(), 0
        elif wait >= patience and epoch > 0:
            st.set_weights(best_w)
            break
    return hist, (ckpt_w if ckpt_w is not None else st.weights())
