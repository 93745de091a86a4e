"""Shared helpers for the -m gpu parity tests (they call the product only through the C-ABI wrappers)."""
import ctypes as C

import numpy as np
import torch

from anime_recommendations_b200 import _capi
from anime_recommendations_b200._capi import ArPlan, ArTable, check, lib, ptr, stream_ptr

DEV = "cuda:0"


def dev(a, dtype=None):
    a = np.ascontiguousarray(a, dtype=dtype)
    return torch.from_numpy(a).to(DEV)


def make_plan(n_slots, batch):
    hc = batch // _capi.AR_HEAVY_LEN + 1
    bufs = dict(order=torch.full((n_slots, batch), -7, dtype=torch.int32, device=DEV),
                uniq=torch.full((n_slots, batch), -7, dtype=torch.int32, device=DEV),
                off=torch.full((n_slots, batch + 1), -7, dtype=torch.int32, device=DEV),
                meta=torch.zeros((n_slots, 4), dtype=torch.int32, device=DEV),
                heavy=torch.full((n_slots, hc), -7, dtype=torch.int32, device=DEV))
    p = ArPlan()
    p.batch_cap, p.heavy_cap, p.n_slots = batch, hc, n_slots
    for k, v in bufs.items():
        setattr(p, k, v.data_ptr())
    return p, bufs


def make_table(W):
    W = dev(W, np.float32)
    m, v = torch.zeros_like(W), torch.zeros_like(W)
    last = torch.zeros(W.shape[0], dtype=torch.int32, device=DEV)
    t = ArTable()
    t.n_rows, t.dim = W.shape
    t.W, t.m, t.v, t.last_step = W.data_ptr(), m.data_ptr(), v.data_ptr(), last.data_ptr()
    return t, dict(W=W, m=m, v=v, last=last)


def assert_topk_close(idx, sc, oidx, osc, full_scores=None, tol=3e-6):
    """Index lists equal except where the oracle's own scores tie within `tol` (documented tie rule)."""
    idx, sc, oidx, osc = map(np.asarray, (idx, sc, oidx, osc))
    assert idx.shape == oidx.shape, (idx, oidx)
    np.testing.assert_allclose(sc, osc, rtol=0, atol=tol)
    for j in np.nonzero(idx != oidx)[0]:
        assert full_scores is not None, (j, idx, oidx)
        assert abs(float(full_scores[idx[j]]) - float(full_scores[oidx[j]])) <= tol, (j, idx[j], oidx[j])
