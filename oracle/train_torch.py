"""Multi-threaded CPU port of oracle/train.py::train_step (TEST/BASELINE INFRASTRUCTURE).

Same arithmetic as the NumPy oracle (dense gradient = scatter-add + 2*l2*W, dense Keras Adam over
both tables every step -- what TensorFlow 2.12 executes for neural_network.py:66-106), written
with PyTorch CPU ops so that it uses every host core.  Used only as the `cpu_baseline` /
`--impl reference` leg of bench.py ("CPU restatement of the reference's TF path; TensorFlow is
not installable in this image") and cross-checked against the NumPy oracle in tests.
"""
from __future__ import annotations

import torch

from . import train as _t


class TorchState:
    def __init__(self, st: "_t.State"):
        f = torch.float32
        self.U, self.A = torch.from_numpy(st.U.copy()), torch.from_numpy(st.A.copy())
        self.mU, self.vU = torch.from_numpy(st.mU.copy()), torch.from_numpy(st.vU.copy())
        self.mA, self.vA = torch.from_numpy(st.mA.copy()), torch.from_numpy(st.vA.copy())
        self.head = torch.tensor(st.head, dtype=f)
        self.mh, self.vh = torch.tensor(st.mh, dtype=f), torch.tensor(st.vh, dtype=f)
        self.mov = torch.tensor([float(st.mov_mean), float(st.mov_var)], dtype=f)
        self.iterations = st.iterations


def _adam(theta, m, v, g, alpha):
    m.add_((g - m) * (1 - _t.BETA1))
    v.add_((g * g - v) * (1 - _t.BETA2))
    theta.sub_((m * alpha) / (v.sqrt() + _t.ADAM_EPS))


@torch.no_grad()
def train_step(s: TorchState, iu, ia, t, lr, l2=1e-4):
    iu = torch.as_tensor(iu, dtype=torch.int64)
    ia = torch.as_tensor(ia, dtype=torch.int64)
    t = torch.as_tensor(t, dtype=torch.float32)
    n = iu.numel()
    u, a = s.U[iu], s.A[ia]
    ru = torch.rsqrt(torch.clamp((u * u).sum(1), min=_t.L2NORM_EPS))
    ra = torch.rsqrt(torch.clamp((a * a).sum(1), min=_t.L2NORM_EPS))
    uh, ah = u * ru[:, None], a * ra[:, None]
    c = (uh * ah).sum(1)
    w, b, gamma, beta = s.head
    z = w * c + b
    mu = z.mean()
    var = ((z - mu) ** 2).mean()
    inv = torch.rsqrt(var + _t.BN_EPS)
    zh = (z - mu) * inv
    y = gamma * zh + beta
    p = torch.sigmoid(y)
    bce = (torch.clamp(y, min=0) - y * t + torch.log1p(torch.exp(-y.abs()))).mean()
    mse = ((t - p) ** 2).mean()
    reg = l2 * ((s.U * s.U).sum() + (s.A * s.A).sum())
    dy = (p - t) / n
    dgamma, dbeta = (dy * zh).sum(), dy.sum()
    dzh = gamma * dy
    dz = inv / n * (n * dzh - dzh.sum() - zh * (dzh * zh).sum())
    ghead = torch.stack([(dz * c).sum(), dz.sum(), dgamma, dbeta])
    dc = w * dz
    du = (ru * dc)[:, None] * (ah - c[:, None] * uh)
    da = (ra * dc)[:, None] * (uh - c[:, None] * ah)
    gU = (2 * l2) * s.U
    gA = (2 * l2) * s.A
    gU.index_add_(0, iu, du)
    gA.index_add_(0, ia, da)
    step = s.iterations + 1
    alpha = float(_t.adam_alpha(lr, step))
    _adam(s.U, s.mU, s.vU, gU, alpha)
    _adam(s.A, s.mA, s.vA, gA, alpha)
    _adam(s.head, s.mh, s.vh, ghead, alpha)
    s.mov[0] -= (s.mov[0] - mu) * (1 - _t.BN_MOMENTUM)
    s.mov[1] -= (s.mov[1] - var) * (1 - _t.BN_MOMENTUM)
    s.iterations = step
    return dict(bce=float(bce), reg=float(reg), loss=float(bce + reg), mse=float(mse), n=n)
