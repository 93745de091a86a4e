"""Keras-like facade over the CUDA training path (host-side mirror of the reference interface).

Mirrors the object surface the reference's components use:
`Model.fit(x=[users, animes], y, batch_size, epochs, verbose, validation_data, callbacks)` ->
`.history` (neural_network.py:210-217,223), `Model.save(path)` (:221), `load_model(path)`
(similar_anime.py:132), `Model.get_layer(name).get_weights()[0]` (similar_anime.py:155-157),
`Model.predict([user_arr, anime_arr])` -> (M,1) float32 (model_recs.py:394).

All arithmetic runs in libanimerec.so (hand-written sm_100a CUDA); torch only owns the device
buffers.  There is no CPU fallback: constructing a model without a CUDA device raises.
"""
from __future__ import annotations

import ctypes as C
import math
import os
import time

import numpy as np
import torch

from . import _capi
from ._capi import ArPlan, ArSched, ArTable, ArTrainCtx, check, lib, ptr, stream_ptr
from . import weights_io

BETA1, BETA2 = 0.9, 0.999
PLAN_CHUNK = 256
REPLAY_DEPTH = int(os.environ.get("AR_REPLAY_DEPTH", "3"))   # look-ahead depth of the replay schedule


def adam_alpha_table(lr, t_first, count):
    """alpha_t = lr*sqrt(1-b2^t)/(1-b1^t) in float32 for t = t_first .. t_first+count-1 (Keras Adam)."""
    t = np.arange(t_first, t_first + count, dtype=np.float32)
    b1p = np.power(np.float32(BETA1), t)
    b2p = np.power(np.float32(BETA2), t)
    return (np.float32(lr) * np.sqrt(np.float32(1) - b2p) / (np.float32(1) - b1p)).astype(np.float32)


def he_normal_scalar(rng):
    """Dense(1) kernel (1,1) under kernel_initializer='he_normal' (config.yaml:56)."""
    sd = math.sqrt(2.0) / 0.87962566103423978
    while True:
        x = rng.standard_normal()
        if abs(x) <= 2.0:
            return np.float32(x * sd)


class History:
    def __init__(self):
        self.history = {}
        self.epoch = []


class _EmbeddingLayer:
    def __init__(self, model, which, name):
        self._model, self._which, self.name = model, which, name

    def get_weights(self):
        self._model._sync_tables()
        t = self._model.U if self._which == "user" else self._model.A
        return [t.detach().cpu().numpy()]


class _ScalarLayer:
    def __init__(self, name, getter):
        self.name, self._getter = name, getter

    def get_weights(self):
        return self._getter()


class EmbeddingDotModel:
    """The model of neural_network.py:66-106 with its optimizer state, resident in HBM."""

    def __init__(self, n_users, n_anime, embedding_size=128, l2_reg_factor=1e-4,
                 kernel_initializer="he_normal", ID_emb_name="user_embedding",
                 anime_emb_name="anime_embedding", merged_name="dot_product", seed=None,
                 device=None, adam_mode="replay", dense_kernel=None, device_init=False):
        if not torch.cuda.is_available():
            raise _capi.AnimerecError("EmbeddingDotModel needs a CUDA device (sm_100a); there is no CPU fallback")
        lib()
        check(lib().ar_check_device(), "ar_check_device")
        if adam_mode not in _capi.ADAM_MODES:
            raise ValueError("adam_mode must be one of %s" % list(_capi.ADAM_MODES))
        D = int(embedding_size)
        if D % 4 or not 0 < D <= 512:
            raise ValueError("embedding_size must be a multiple of 4 in (0, 512]")
        self.device = torch.device(device or "cuda:%d" % torch.cuda.current_device())
        self.n_users, self.n_anime, self.dim = int(n_users), int(n_anime), D
        self.l2 = float(l2_reg_factor)
        self.adam_mode = adam_mode
        self.names = dict(user=ID_emb_name, anime=anime_emb_name, merged=merged_name)
        rng = np.random.RandomState(seed)
        if kernel_initializer != "he_normal" and dense_kernel is None:
            raise ValueError("only kernel_initializer='he_normal' (config.yaml:56) or an explicit dense_kernel")
        w = he_normal_scalar(rng) if dense_kernel is None else np.float32(dense_kernel)
        if device_init:    # very large tables (cfg5: 10 M x 256): draw U(-0.05, 0.05) on the device
            g = torch.Generator(device=self.device)
            g.manual_seed(0 if seed is None else int(seed))
            U = torch.rand((self.n_users, D), generator=g, device=self.device) * 0.1 - 0.05
            A = torch.rand((self.n_anime, D), generator=g, device=self.device) * 0.1 - 0.05
        else:
            U = rng.uniform(-0.05, 0.05, size=(self.n_users, D)).astype(np.float32)
            A = rng.uniform(-0.05, 0.05, size=(self.n_anime, D)).astype(np.float32)
        self._alloc(U, A, np.array([w, 0.0, 1.0, 0.0], np.float32), np.array([0.0, 1.0], np.float32))
        self.iterations = 0
        self.lr = 1e-3                      # Keras Adam default until a LearningRateScheduler sets it
        self.stop_training = False
        self.history = None
        self._alpha = None                  # device alpha table, index = global step
        self._alpha_host = np.zeros(1, np.float32)
        self._stepw = None                  # device step-weight table (samples in step / batch), same indexing
        self.timings = {}

    # ------------------------------------------------------------------ state
    def _alloc(self, U, A, head, bn):
        dev = self.device
        f = dict(dtype=torch.float32, device=dev)
        as_dev = lambda t: t.to(dev).contiguous() if isinstance(t, torch.Tensor) else \
            torch.from_numpy(np.ascontiguousarray(t, np.float32)).to(dev)  # noqa: E731
        self.U, self.A = as_dev(U), as_dev(A)
        self.mU, self.vU = torch.zeros_like(self.U), torch.zeros_like(self.U)
        self.mA, self.vA = torch.zeros_like(self.A), torch.zeros_like(self.A)
        self.lastU = torch.zeros(self.n_users, dtype=torch.int32, device=dev)
        self.lastA = torch.zeros(self.n_anime, dtype=torch.int32, device=dev)
        self.head = torch.from_numpy(np.asarray(head, np.float32)).to(dev)
        self.head_m = torch.zeros(4, **f)
        self.head_v = torch.zeros(4, **f)
        self.bn_moving = torch.from_numpy(np.asarray(bn, np.float32)).to(dev)
        # replay-schedule state (ar_plan_sched): global step of every row's latest planned touch, and the step
        # of the last full flush
        self.seenU = torch.zeros(self.n_users, dtype=torch.int32, device=dev)
        self.seenA = torch.zeros(self.n_anime, dtype=torch.int32, device=dev)
        self._t_flush = 0
        # L2-regulariser accumulator of the reported loss (fixed point, see animerec.h ar_train_ctx.reg_acc)
        self.reg_acc = torch.zeros(1, dtype=torch.int64, device=dev)
        self.reg_scale = 1.0
        self._reg_on = False

    def _table(self, which):
        t = ArTable()
        if which == "user":
            t.n_rows, t.dim = self.n_users, self.dim
            t.W, t.m, t.v, t.last_step = (x.data_ptr() for x in (self.U, self.mU, self.vU, self.lastU))
        else:
            t.n_rows, t.dim = self.n_anime, self.dim
            t.W, t.m, t.v, t.last_step = (x.data_ptr() for x in (self.A, self.mA, self.vA, self.lastA))
        return t

    def _ensure_alpha(self, upto):
        """Device alpha table covering global steps [0, upto]."""
        if self._alpha is None or self._alpha.numel() <= upto:
            n = max(upto + 1, 2 * (0 if self._alpha is None else self._alpha.numel()))
            host = np.zeros(n, np.float32)
            host[:len(self._alpha_host)] = self._alpha_host
            self._alpha_host = host
            self._alpha = torch.from_numpy(host).to(self.device)
            sw = torch.ones(n, dtype=torch.float32, device=self.device)
            if self._stepw is not None:
                sw[:self._stepw.numel()].copy_(self._stepw)
            self._stepw = sw

    def _set_alpha(self, lr, t_first, count):
        self._ensure_alpha(t_first + count)
        vals = adam_alpha_table(lr, t_first, count)
        self._alpha_host[t_first:t_first + count] = vals
        self._alpha[t_first:t_first + count].copy_(torch.from_numpy(vals), non_blocking=False)

    def _sync_tables(self):
        """In replay mode rows lag behind; replay every row up to the current optimizer step."""
        if self.adam_mode != "replay" or self.iterations == 0 or self._t_flush == self.iterations:
            return
        self._ensure_alpha(self.iterations)
        st = stream_ptr()
        reg = (ptr(self.reg_acc), ptr(self._stepw), self.reg_scale) if self._reg_on else (None, None, 1.0)
        for which in ("user", "anime"):
            t = self._table(which)
            check(lib().ar_table_flush(C.byref(t), ptr(self._alpha), self.l2, self.iterations, *reg, st), "ar_table_flush")
        self._t_flush = self.iterations

    def _begin_reg(self, steps, batch, n_last):
        """Arm the regulariser accumulator for an epoch of `steps` steps starting at the current step."""
        t0 = self.iterations
        self._ensure_alpha(t0 + steps)
        self._stepw[t0 + 1:t0 + steps + 1] = 1.0
        if n_last != batch:
            self._stepw[t0 + steps] = float(n_last) / float(batch)
        ss = self.reg_sumsq()
        bound = max(1.0, float(steps) * (4.0 * ss + 1.0))
        self.reg_scale = float(2.0 ** min(40, int(math.floor(61 - math.log2(bound)))))
        self.reg_acc.zero_()
        self._reg_on = True
        return ss

    def _end_reg(self):
        """sum over the epoch's steps t of stepw[t] * (sum U_{t-1}^2 + sum A_{t-1}^2); tables are flushed first."""
        self._sync_tables()
        self._reg_on = False
        return float(self.reg_acc.item()) / self.reg_scale

    def reg_sumsq(self):
        """sum U^2 + sum A^2 of the current (synchronised) tables, float64 on host."""
        self._sync_tables()
        out = torch.zeros(1, dtype=torch.float64, device=self.device)
        st = stream_ptr()
        check(lib().ar_sumsq(ptr(self.U), self.U.numel(), ptr(out), st), "ar_sumsq")
        check(lib().ar_sumsq(ptr(self.A), self.A.numel(), ptr(out), st), "ar_sumsq")
        return float(out.item())

    # ------------------------------------------------------------------ Keras surface
    def get_layer(self, name):
        if name == self.names["user"]:
            return _EmbeddingLayer(self, "user", name)
        if name == self.names["anime"]:
            return _EmbeddingLayer(self, "anime", name)
        if name == "dense":
            return _ScalarLayer(name, lambda: [self.head[0:1].cpu().numpy().reshape(1, 1), self.head[1:2].cpu().numpy()])
        if name == "batch_normalization":
            return _ScalarLayer(name, lambda: [self.head[2:3].cpu().numpy(), self.head[3:4].cpu().numpy(),
                                               self.bn_moving[0:1].cpu().numpy(), self.bn_moving[1:2].cpu().numpy()])
        raise ValueError("No such layer: %s. Existing layers are: %s" % (
            name, [self.names["user"], self.names["anime"], self.names["merged"], "dense", "batch_normalization"]))

    def get_weights(self):
        """[user table, anime table, dense kernel, dense bias, gamma, beta, moving_mean, moving_variance]."""
        self._sync_tables()
        h, b = self.head.cpu().numpy(), self.bn_moving.cpu().numpy()
        return [self.U.cpu().numpy(), self.A.cpu().numpy(), h[0:1].reshape(1, 1), h[1:2], h[2:3], h[3:4],
                b[0:1], b[1:2]]

    def set_weights(self, w):
        self._sync_tables()
        self.U.copy_(torch.from_numpy(np.ascontiguousarray(w[0], np.float32)))
        self.A.copy_(torch.from_numpy(np.ascontiguousarray(w[1], np.float32)))
        head = np.array([np.ravel(w[2])[0], np.ravel(w[3])[0], np.ravel(w[4])[0], np.ravel(w[5])[0]], np.float32)
        self.head.copy_(torch.from_numpy(head))
        self.bn_moving.copy_(torch.from_numpy(np.array([np.ravel(w[6])[0], np.ravel(w[7])[0]], np.float32)))

    def _device_weights(self):
        self._sync_tables()
        return [self.U.clone(), self.A.clone(), self.head.clone(), self.bn_moving.clone()]

    def _restore_device_weights(self, w):
        self._sync_tables()
        self.U.copy_(w[0]); self.A.copy_(w[1]); self.head.copy_(w[2]); self.bn_moving.copy_(w[3])

    @staticmethod
    def _as_idx(a, dev):
        if isinstance(a, torch.Tensor):
            return a.to(device=dev, dtype=torch.int32).reshape(-1).contiguous()
        a = np.asarray(a).reshape(-1)
        if a.dtype.kind == "f":          # Keras Input(shape=[1]) carries ids as float32
            a = a.astype(np.int64)
        return torch.from_numpy(np.ascontiguousarray(a, np.int32)).to(dev)

    @staticmethod
    def _as_f32(a, dev):
        if isinstance(a, torch.Tensor):
            return a.to(device=dev, dtype=torch.float32).reshape(-1).contiguous()
        return torch.from_numpy(np.ascontiguousarray(np.asarray(a, np.float64).reshape(-1), np.float32)).to(dev)

    def predict(self, x, verbose=0, batch_size=None):
        """model.predict([user_arr, anime_arr]) -> (M,1) float32, inference-mode BatchNorm."""
        self._sync_tables()
        iu, ia = self._as_idx(x[0], self.device), self._as_idx(x[1], self.device)
        if iu.numel() != ia.numel():
            raise ValueError("user and anime index arrays differ in length")
        self._check_range(iu, ia)
        out = torch.empty(iu.numel(), dtype=torch.float32, device=self.device)
        check(lib().ar_predict(ptr(self.U), ptr(self.A), self.dim, ptr(self.head), ptr(self.bn_moving),
                               ptr(iu), ptr(ia), iu.numel(), ptr(out), stream_ptr()), "ar_predict")
        return out.cpu().numpy().reshape(-1, 1)

    def _check_range(self, iu, ia):
        if iu.numel() and (int(iu.min()) < 0 or int(iu.max()) >= self.n_users):
            raise IndexError("user index outside [0, %d)" % self.n_users)
        if ia.numel() and (int(ia.min()) < 0 or int(ia.max()) >= self.n_anime):
            raise IndexError("anime index outside [0, %d)" % self.n_anime)

    def evaluate(self, x, y, batch_size=None, verbose=0, return_dict=True):
        """Keras test pass: inference BN, loss = mean BCE + l2*(sum U^2 + sum A^2), mse."""
        self._sync_tables()
        iu, ia = self._as_idx(x[0], self.device), self._as_idx(x[1], self.device)
        t = self._as_f32(y, self.device)
        self._check_range(iu, ia)
        sums = torch.zeros(2, dtype=torch.float64, device=self.device)
        check(lib().ar_eval_sums(ptr(self.U), ptr(self.A), self.dim, ptr(self.head), ptr(self.bn_moving),
                                 ptr(iu), ptr(ia), ptr(t), iu.numel(), ptr(sums), stream_ptr()), "ar_eval_sums")
        s = sums.cpu().numpy()
        n = max(1, iu.numel())
        reg = self.l2 * self.reg_sumsq()
        out = dict(loss=s[0] / n + reg, mse=s[1] / n, bce=s[0] / n, reg=reg)
        return out if return_dict else [out["loss"], out["mse"]]

    # ------------------------------------------------------------------ training
    def fit(self, x, y, batch_size=10000, epochs=1, verbose=0, validation_data=None, callbacks=None,
            shuffle="numpy", shuffle_seed=0, initial_epoch=0):
        """model.fit of neural_network.py:210-217.

        shuffle: "numpy"  -> np.random.RandomState(shuffle_seed + epoch).permutation(n) (the oracle's
                             documented rule; Keras' own shuffle is unseeded),
                 "device" -> torch.randperm on the GPU (fast path for 1e8 samples),
                 False    -> keep the given order.
        """
        dev = self.device
        iu_all, ia_all = self._as_idx(x[0], dev), self._as_idx(x[1], dev)
        y_all = self._as_f32(y, dev)
        N = iu_all.numel()
        if not (ia_all.numel() == N == y_all.numel()) or N == 0:
            raise ValueError("x[0], x[1] and y must be non-empty and of equal length")
        B = int(batch_size)
        if not 0 < B <= _capi.AR_MAX_BATCH:
            raise ValueError("batch_size must be in (0, %d]" % _capi.AR_MAX_BATCH)
        self._check_range(iu_all, ia_all)
        steps = (N + B - 1) // B
        sess = TrainSession(self, B, total_steps=max(0, epochs - initial_epoch) * steps)

        callbacks = list(callbacks or [])
        for cb in callbacks:
            cb.set_model(self)
        hist = History()
        self.history = hist
        self.stop_training = False
        for cb in callbacks:
            cb.on_train_begin()
        val = None
        if validation_data is not None:
            (vx, vy) = validation_data[0], validation_data[1]
            val = (self._as_idx(vx[0], dev), self._as_idx(vx[1], dev), self._as_f32(vy, dev))
            self._check_range(val[0], val[1])
        for epoch in range(initial_epoch, epochs):
            t_epoch = time.perf_counter()
            for cb in callbacks:
                cb.on_epoch_begin(epoch)
            lr = float(self.lr)
            t0 = self.iterations
            if shuffle == "numpy":
                perm = torch.from_numpy(np.random.RandomState(shuffle_seed + epoch).permutation(N)).to(dev)
            elif shuffle == "device":
                g = torch.Generator(device=dev)
                g.manual_seed(shuffle_seed + epoch)
                perm = torch.randperm(N, device=dev, generator=g)
            elif shuffle in (False, None, "none"):
                perm = None
            else:
                raise ValueError("shuffle must be 'numpy', 'device' or False")
            if perm is None:
                iu_e, ia_e, y_e = iu_all, ia_all, y_all
            else:
                iu_e, ia_e, y_e = iu_all[perm].contiguous(), ia_all[perm].contiguous(), y_all[perm].contiguous()
            dbg = os.environ.get("AR_FIT_TIMING") == "1"
            tick = []

            def _tk(name):
                if dbg:
                    torch.cuda.synchronize()
                    tick.append((name, time.perf_counter()))
            _tk("start")
            reg0 = self.l2 * self._begin_reg(steps, B, N - (steps - 1) * B)
            _tk("reg0")
            sess.run(iu_e, ia_e, y_e, lr)
            _tk("steps")
            acc = self._end_reg()          # flushes the tables, then reads the accumulator
            sess.check_health()
            _tk("flush")

            m = sess.metrics[t0 + 1:t0 + steps + 1].cpu().numpy().astype(np.float64)
            w = m[:, 2]
            bce = float((m[:, 0] * w).sum() / N)
            mse = float((m[:, 1] * w).sum() / N)
            reg1 = self.l2 * self.reg_sumsq()
            if self.adam_mode == "touched":
                reg = 0.5 * (reg0 + reg1)   # not the reference's arithmetic anyway: end-point average
            else:
                # Keras: loss_t = BCE_t + l2*(sum U^2 + sum A^2) with the weights BEFORE step t, averaged with the
                # step's sample count as weight (neural_network.py:73,78,85; oracle fit())
                reg = self.l2 * acc * B / N
            logs = dict(loss=bce + reg, mse=mse)
            if val is not None:
                sums = torch.zeros(2, dtype=torch.float64, device=dev)
                check(lib().ar_eval_sums(ptr(self.U), ptr(self.A), self.dim, ptr(self.head), ptr(self.bn_moving),
                                         ptr(val[0]), ptr(val[1]), ptr(val[2]), val[0].numel(), ptr(sums),
                                         stream_ptr()), "ar_eval_sums")
                sv = sums.cpu().numpy()
                nv = max(1, val[0].numel())
                logs["val_loss"] = float(sv[0] / nv + reg1)
                logs["val_mse"] = float(sv[1] / nv)
            logs["lr"] = float(np.float32(lr))
            _tk("metrics+val")
            if dbg:
                self.timings.setdefault("sections", []).append(
                    {b[0]: b[1] - a[1] for a, b in zip(tick[:-1], tick[1:])})
            self.timings.setdefault("epoch_s", []).append(time.perf_counter() - t_epoch)
            self.last_epoch_parts = dict(bce=bce, reg=reg, reg_end=reg1)
            for k, v in logs.items():
                hist.history.setdefault(k, []).append(v)
            hist.epoch.append(epoch)
            if verbose:
                print("Epoch %d/%d - %.2fs - %s" % (epoch + 1, epochs, self.timings["epoch_s"][-1],
                                                     " - ".join("%s: %.6g" % kv for kv in logs.items())))
            for cb in callbacks:
                cb.on_epoch_end(epoch, logs)
            if self.stop_training:
                break
        for cb in callbacks:
            cb.on_train_end()
        return hist

    # ------------------------------------------------------------------ persistence
    def save(self, path, include_optimizer=True):
        """model.save(path): full model in the Keras-2.12 H5 name layout (SURVEY §5 checkpoint row)."""
        weights_io.save_model(self, path, include_optimizer=include_optimizer)

    def save_weights(self, path):
        weights_io.save_model(self, path, include_optimizer=False, weights_only=True)

    def load_weights(self, path):
        weights_io.load_into(self, path)


class TrainSession:
    """Device-side buffers of one training run: two sets of dedup plans + replay schedules (PLAN_CHUNK steps at a
    time; chunk i+1 is planned on a side stream while chunk i trains), the per-step scratch, the per-step
    metrics and the ar_train_ctx handed to libanimerec."""

    def __init__(self, model, batch, total_steps, plan_cap=None):
        self.model, self.B = model, int(batch)
        dev, D, B = model.device, model.dim, int(batch)
        # plan_cap: per-step capacity of the plans and the per-sample scratch (peer mode lists up to plan_cap
        # samples of the GLOBAL batch per rank and step); the per-rank batch otherwise
        P = self.P = int(plan_cap or B)
        self.n_slots = max(1, min(int(total_steps) if total_steps else PLAN_CHUNK, PLAN_CHUNK))
        self.plan_u, self._keep_u = self._make_plan(self.n_slots, P, dev)
        self.plan_a, self._keep_a = self._make_plan(self.n_slots, P, dev)
        f = dict(dtype=torch.float32, device=dev)
        self.uh, self.ah = torch.empty((P, D), **f), torch.empty((P, D), **f)
        self.c, self.ru, self.ra, self.dy = (torch.empty(P, **f) for _ in range(4))
        self.fwd_part = torch.zeros(2 * ((P + 7) // 8), dtype=torch.float64, device=dev)
        self.head_part = torch.zeros(8 * ((P + 255) // 256), dtype=torch.float64, device=dev)
        self.stepc = torch.zeros(16, **f)
        self.ticket = torch.zeros(1, dtype=torch.int32, device=dev)
        self.sched_ws = torch.zeros(2 * (3 * 2 * P + 4), dtype=torch.int32, device=dev)
        info = (C.c_int64 * 8)()
        check(lib().ar_chunk_ws_info(self.n_slots, P, D, info), "ar_chunk_ws_info")
        self._ws_stamps, self._ws_stats, self._ws_nstamps = int(info[1]), int(info[2]), int(info[3])
        self.chunk_ws = torch.zeros(int(info[0]) // 8, dtype=torch.int64, device=dev)   # persistent step kernel's workspace
        self.health = torch.zeros(4, dtype=torch.int32, device=dev)
        self.depth = max(1, min(REPLAY_DEPTH, _capi.AR_SCHED_MAX_DEPTH))
        self.t_cap = model.iterations + int(total_steps)
        model._ensure_alpha(self.t_cap)
        self.metrics = torch.zeros((self.t_cap + 1, 4), **f)
        self.launches = 0
        self.enqueue_s = 0.0    # host time spent inside ar_train_steps* (queueing the launches)
        self._sets = None       # single-GPU path: [set 0, set 1] of (plans, schedule, events), made on first run()
        self.plan_stream = None

    @staticmethod
    def _make_plan(n_slots, batch, dev):
        hc = batch // _capi.AR_HEAVY_LEN + 1
        bufs = dict(order=torch.empty((n_slots, batch), dtype=torch.int32, device=dev),
                    uniq=torch.empty((n_slots, batch), dtype=torch.int32, device=dev),
                    off=torch.empty((n_slots, batch + 1), dtype=torch.int32, device=dev),
                    meta=torch.zeros((n_slots, 4), dtype=torch.int32, device=dev),
                    heavy=torch.empty((n_slots, hc), dtype=torch.int32, device=dev),
                    in_prev=torch.zeros((n_slots, batch), dtype=torch.uint8, device=dev))
        p = ArPlan()
        p.batch_cap, p.heavy_cap, p.n_slots = batch, hc, n_slots
        for k, v in bufs.items():
            setattr(p, k, v.data_ptr())
        return p, bufs

    @staticmethod
    def _make_sched(n_slots, batch, dev):
        i32 = dict(dtype=torch.int32, device=dev)
        cap = 4 * batch      # room for the long rows' split items (ar_sched)
        bufs = dict(codes=torch.zeros((n_slots, cap), **i32), glen=torch.zeros((n_slots, cap), **i32),
                    sub=torch.zeros((n_slots, _capi.AR_SCHED_SUB), **i32),
                    cursor=torch.zeros((n_slots, _capi.AR_SCHED_SUB), **i32),
                    gap_u=torch.zeros((n_slots, batch), **i32), gap_a=torch.zeros((n_slots, batch), **i32),
                    bounds=torch.zeros((n_slots, _capi.AR_SCHED_PARTS + 1), **i32))
        sc = ArSched()
        sc.cap, sc.n_slots = cap, n_slots
        for k, v in bufs.items():
            setattr(sc, k, v.data_ptr())
        return sc, bufs

    def _make_sets(self):
        dev = self.model.device
        self._sets = []
        for k in range(2):
            if k == 0:
                pu, ku, pa, ka = self.plan_u, self._keep_u, self.plan_a, self._keep_a
            else:
                pu, ku = self._make_plan(self.n_slots, self.P, dev)
                pa, ka = self._make_plan(self.n_slots, self.P, dev)
            sc, ks = self._make_sched(self.n_slots, self.P, dev)
            self._sets.append(dict(plan_u=pu, plan_a=pa, keep=(ku, ka, ks), sched=sc,
                                   planned=torch.cuda.Event(), consumed=torch.cuda.Event()))
        self.plan_stream = torch.cuda.Stream(device=dev)

    def _ctx(self, iu, ia, y):
        m = self.model
        ctx = ArTrainCtx()
        ctx.users, ctx.anime = m._table("user"), m._table("anime")
        ctx.head, ctx.head_m, ctx.head_v = m.head.data_ptr(), m.head_m.data_ptr(), m.head_v.data_ptr()
        ctx.bn_moving, ctx.alpha = m.bn_moving.data_ptr(), m._alpha.data_ptr()
        ctx.iu, ctx.ia, ctx.label = iu.data_ptr(), ia.data_ptr(), y.data_ptr()
        ctx.n_samples, ctx.batch, ctx.l2 = iu.numel(), self.B, m.l2
        ctx.mode = _capi.ADAM_MODES[m.adam_mode]
        ctx.plan_u, ctx.plan_a = self.plan_u, self.plan_a
        ctx.uh, ctx.ah, ctx.c, ctx.ru, ctx.ra, ctx.dy = (z.data_ptr() for z in (
            self.uh, self.ah, self.c, self.ru, self.ra, self.dy))
        ctx.fwd_part, ctx.head_part = self.fwd_part.data_ptr(), self.head_part.data_ptr()
        ctx.stepc, ctx.ticket = self.stepc.data_ptr(), self.ticket.data_ptr()
        ctx.metrics = self.metrics.data_ptr()
        if m._reg_on:
            ctx.reg_acc, ctx.stepw, ctx.reg_scale = m.reg_acc.data_ptr(), m._stepw.data_ptr(), m.reg_scale
        ctx.sched_ws = self.sched_ws.data_ptr() if os.environ.get("AR_NO_LPT") is None else None
        ctx.depth = self.depth
        ctx.chunk_ws, ctx.health = self.chunk_ws.data_ptr(), self.health.data_ptr()
        return ctx

    def _plan_chunk(self, st, iu, ia, s0, ns, t0):
        """Queue the planning of steps [s0, s0+ns) (global steps t0+s0+1 ..) into set `st` on the CURRENT stream."""
        m, L, sp = self.model, lib(), stream_ptr()
        N = iu.numel()
        check(L.ar_plan_build(ptr(iu), N, self.B, s0, ns, C.byref(st["plan_u"]), sp), "ar_plan_build(users)")
        check(L.ar_plan_build(ptr(ia), N, self.B, s0, ns, C.byref(st["plan_a"]), sp), "ar_plan_build(anime)")
        if m.adam_mode == "replay":
            check(L.ar_plan_sched(C.byref(st["plan_u"]), C.byref(st["plan_a"]), ns, t0 + s0, m._t_flush,
                                  ptr(m.seenU), m.n_users, ptr(m.seenA), m.n_anime, self.depth, m.dim,
                                  C.byref(st["sched"]), sp), "ar_plan_sched")

    def run(self, iu, ia, y, lr):
        """Train on every sample of (iu, ia, y) (device int32/int32/float32, visit order) at learning
        rate `lr`: ceil(n/B) optimizer steps, one persistent kernel per chunk of PLAN_CHUNK steps.
        Asynchronous on the current stream."""
        m, B = self.model, self.B
        N = iu.numel()
        steps = (N + B - 1) // B
        t0 = m.iterations
        if t0 + steps > self.t_cap:
            raise _capi.AnimerecError("TrainSession sized for %d optimizer steps, %d requested" % (
                self.t_cap, t0 + steps))
        m._set_alpha(lr, t0 + 1, steps)
        if self._sets is None:
            self._make_sets()
        ctx = self._ctx(iu, ia, y)
        main, L, S = torch.cuda.current_stream(), lib(), self.n_slots
        plan_launches = 7 if m.adam_mode == "replay" else 2
        chunks = [(s0, min(S, steps - s0)) for s0 in range(0, steps, S)]
        self.plan_stream.wait_stream(main)                 # the inputs (H2D copies, the shuffle) are queued on `main`
        with torch.cuda.stream(self.plan_stream):
            self._plan_chunk(self._sets[0], iu, ia, *chunks[0], t0)
            self._sets[0]["planned"].record()
        for i, (s0, ns) in enumerate(chunks):
            st = self._sets[i % 2]
            if i + 1 < len(chunks):                        # plan the next chunk while this one runs
                nxt = self._sets[(i + 1) % 2]
                with torch.cuda.stream(self.plan_stream):
                    if i >= 1:
                        self.plan_stream.wait_event(nxt["consumed"])   # chunk i-1 no longer reads that set
                    self._plan_chunk(nxt, iu, ia, *chunks[i + 1], t0)
                    nxt["planned"].record()
            main.wait_event(st["planned"])
            ctx.plan_u, ctx.plan_a, ctx.sched = st["plan_u"], st["plan_a"], st["sched"]
            tq = time.perf_counter()
            check(L.ar_train_steps(C.byref(ctx), s0, 0, t0 + s0, ns, stream_ptr(main)), "ar_train_steps")
            self.enqueue_s += time.perf_counter() - tq
            st["consumed"].record(main)
            self.launches += 1 + plan_launches             # the step kernel + the chunk's planning kernels (side stream)
            self._last_ns = ns
        main.wait_stream(self.plan_stream)                 # nothing of this call is left on the side stream
        m.iterations = t0 + steps
        self._last_set = self._sets[(len(chunks) - 1) % 2]
        return steps

    def timeline(self):
        """In-kernel timeline of the LAST chunk run (synchronises): per-step phase durations [us] from CTA 0's
        %globaltimer stamps and the replay warps' statistics (ar_chunk_ws_info)."""
        torch.cuda.synchronize()
        ns, k = self._last_ns, self._ws_nstamps
        ws = self.chunk_ws.cpu().numpy()
        st = ws[self._ws_stamps // 8:self._ws_stamps // 8 + self.n_slots * k].reshape(self.n_slots, k)[:ns].astype(np.float64)
        stats = ws[self._ws_stats // 8:self._ws_stats // 8 + 8]
        d = lambda a, b: (st[:, b] - st[:, a]) * 1e-3
        out = dict(steps=ns, gate_us=d(0, 1), fwd_us=d(1, 2), head_us=d(2, 3), update_us=d(3, 4), barrier2_us=d(4, 5),
                   step_us=np.r_[np.diff(st[:, 0]) * 1e-3, (st[-1, 5] - st[-1, 0]) * 1e-3],
                   replay_busy_cycles=int(stats[0]), replay_items=int(stats[1]), replay_element_steps=int(stats[2]),
                   kernel_ns=int(stats[3]), replay_warps=int(stats[4]), kernel_cycles=int(stats[5]))
        if getattr(self, "persistent", False):       # peer kernel: stamp 7 = CTA 0's own forward done, the rest of fwd_us is waiting
            out["fwd_own_us"] = d(1, 7)
        if self.model.adam_mode == "dense":
            out["dense_us"] = d(5, 6)
            out["step_us"] = np.r_[np.diff(st[:, 0]) * 1e-3, (st[-1, 6] - st[-1, 0]) * 1e-3]
        return out

    def check_health(self):
        """Raise if a wait inside the step kernel timed out, or a row update found a row behind schedule (replay
        schedule and plans out of step)."""
        h = self.health.cpu().tolist()
        if h[1]:
            self.health.zero_()
            raise _capi.AnimerecError("the step kernel aborted: %d CTAs gave up waiting; results of this run are "
                                      "invalid" % h[1])
        if h[0] and self.model.adam_mode == "replay":
            self.health.zero_()
            raise _capi.AnimerecError("%d row updates found their row behind the replay schedule" % h[0])


def load_model(path, device=None, adam_mode="replay"):
    """tf.keras.models.load_model replacement (similar_anime.py:132, model_recs.py:204)."""
    return weights_io.load_model(path, device=device, adam_mode=adam_mode)
